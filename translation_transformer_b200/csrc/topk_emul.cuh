// torch-CPU `topk(1)` tie-break emulation shared by the greedy and beam-search accept kernels.
#pragma once
#include "common.cuh"

namespace ttb {

// ---- torch CPU topk(1) tie-break: libstdc++ introselect on (value, index) pairs -------------
struct VI { int v; int i; };
__device__ __forceinline__ bool gt(const VI& a, const VI& b) { return a.v > b.v; }
// (value << 8 | index) in one word: half the memory operations of VI; for n < 64 and small values
struct PK { int x; };
__device__ __forceinline__ bool gt(const PK& a, const PK& b) { return (a.x >> 8) > (b.x >> 8); }
template <typename VI> __device__ __forceinline__ void swp(VI& a, VI& b) { VI t = a; a = b; b = t; }

template <typename VI> __device__ inline void push_heap_(VI* e, int first, int hole, int top, VI val) {
    int parent = (hole - 1) / 2;
    while (hole > top && gt(e[first + parent], val)) {
        e[first + hole] = e[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    e[first + hole] = val;
}
template <typename VI> __device__ inline void adjust_heap_(VI* e, int first, int hole, int len, VI val) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (gt(e[first + child], e[first + child - 1])) child--;
        e[first + hole] = e[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        e[first + hole] = e[first + child - 1];
        hole = child - 1;
    }
    push_heap_(e, first, hole, top, val);
}
template <typename VI> __device__ inline void heap_select_(VI* e, int first, int middle, int last) {
    const int len = middle - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            adjust_heap_(e, first, parent, len, e[first + parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    for (int i = middle; i < last; ++i) {
        if (gt(e[i], e[first])) {
            VI val = e[i];
            e[i] = e[first];
            adjust_heap_(e, first, 0, len, val);
        }
    }
}
// libstdc++ std::nth_element(e, e + nth, e + n, gt) (introselect)
template <typename VI>
__device__ inline void nth_element_(VI* e, int n, int nth) {
    int first = 0, last = n;
    if (n == 0 || nth == n) return;
    int depth = 2 * (31 - __clz(n));
    while (last - first > 3) {
        if (depth == 0) {
            heap_select_(e, first, nth + 1, last);
            swp(e[first], e[nth]);
            return;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        {   // __move_median_to_first(first, first + 1, mid, last - 1)
            const int a = first + 1, b = mid, c = last - 1;
            if (gt(e[a], e[b])) {
                if (gt(e[b], e[c])) swp(e[first], e[b]);
                else if (gt(e[a], e[c])) swp(e[first], e[c]);
                else swp(e[first], e[a]);
            } else if (gt(e[a], e[c])) swp(e[first], e[a]);
            else if (gt(e[b], e[c])) swp(e[first], e[c]);
            else swp(e[first], e[b]);
        }
        int lo = first + 1, hi = last;  // __unguarded_partition(first + 1, last, pivot = first)
        while (true) {
            while (gt(e[lo], e[first])) ++lo;
            --hi;
            while (gt(e[first], e[hi])) --hi;
            if (!(lo < hi)) break;
            swp(e[lo], e[hi]);
            ++lo;
        }
        if (lo <= nth) first = lo; else last = lo;
    }
    for (int i = first + 1; i < last; ++i) {  // __insertion_sort
        VI val = e[i];
        if (gt(val, e[first])) {
            for (int j = i; j > first; --j) e[j] = e[j - 1];
            e[first] = val;
        } else {
            int j = i;
            while (gt(val, e[j - 1])) { e[j] = e[j - 1]; --j; }
            e[j] = val;
        }
    }
}

template <typename VI> __device__ inline void nth_element0_(VI* e, int n) { nth_element_(e, n, 0); }

// (float value, index) pairs for torch.topk over scores
struct VF { float v; int i; };
__device__ __forceinline__ bool gt(const VF& a, const VF& b) { return a.v > b.v; }

template <typename VI> __device__ inline void insertion_sort_(VI* e, int first, int last) {
    for (int i = first + 1; i < last; ++i) {
        VI val = e[i];
        if (gt(val, e[first])) {
            for (int j = i; j > first; --j) e[j] = e[j - 1];
            e[first] = val;
        } else {
            int j = i;
            while (gt(val, e[j - 1])) { e[j] = e[j - 1]; --j; }
            e[j] = val;
        }
    }
}
// std::partial_sort(e, e + k, e + n, gt): __heap_select + __sort_heap.  The scan over [k, n) skips runs of eight
// entries that cannot enter the heap with independent loads (the heap top only changes when one does).
template <typename VI> __device__ inline void partial_sort_(VI* e, int n, int k) {
    if (k >= 2) {
        int parent = (k - 2) / 2;
        while (true) {
            adjust_heap_(e, 0, parent, k, e[parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    int i = k;
    while (i < n) {
        if (i + 8 <= n) {
            const VI top = e[0];
            bool any = false;
#pragma unroll
            for (int u = 0; u < 8; ++u) any |= gt(e[i + u], top);
            if (!any) { i += 8; continue; }
        }
        if (gt(e[i], e[0])) {
            VI val = e[i];
            e[i] = e[0];
            adjust_heap_(e, 0, 0, k, val);
        }
        ++i;
    }
    int last = k;
    while (last > 1) {   // __sort_heap
        --last;
        VI val = e[last];
        e[last] = e[0];
        adjust_heap_(e, 0, 0, last, val);
    }
}
// The first k entries of e become what torch.topk(values, k, largest=True, sorted=True) returns on the CPU backend
// (ATen/native/cpu/TopKImpl.h: partial_sort when k * 64 <= n, else nth_element + sort of the first k - 1 entries;
// the sort is an insertion sort up to 16 entries, which is also used (inexact among ties) beyond that).
template <typename VI> __device__ inline void topk_sorted_torch_cpu_(VI* e, int n, int k) {
    if ((long long)k * 64 <= n) {
        partial_sort_(e, n, k);
    } else {
        nth_element_(e, n, k - 1);
        if (k - 1 > 1) insertion_sort_(e, 0, k - 1);
    }
}

// index torch.topk(vals, 1) returns on the CPU backend (n < 64: std::nth_element, else partial_sort)
__device__ inline int topk1_torch_cpu(const int* vals, int n) {
    if (n >= 64 || n <= 1) {
        int best = 0;
        for (int j = 1; j < n; ++j) if (vals[j] > vals[best]) best = j;
        return best;
    }
    VI e[64];
    for (int j = 0; j < n; ++j) { e[j].v = vals[j]; e[j].i = j; }
    nth_element0_(e, n);
    return e[0].i;
}
// same on a shared-memory array that already holds (value << 8 | index) words, 1 < n < 64 (clobbers it)
__device__ inline int topk1_torch_cpu_packed(int* packed, int n) {
    nth_element0_(reinterpret_cast<PK*>(packed), n);
    return packed[0] & 0xff;
}



}  // namespace ttb
