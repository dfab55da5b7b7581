"""Oracle (test infrastructure): which index does `torch.topk` return among ties?

The reference picks "the best draft" with `tensor.topk(1, -1)`
(/root/reference/src/decoding/speculative_decoding.py:133, :553, and through
`topk_in_each_group` :223/:232).  Accepted-token counts are small integers, so
ties are the norm and the returned index is decided by the backend's selection
routine, not by the maths.  On the CPU backend (the one the golden vectors were
produced with, torch 2.11) `topk` over a row of length n does

    n >= 64*k : std::partial_sort(first, first+k, last, greater)
    otherwise : std::nth_element(first, first+k-1, last, greater)
                [+ std::sort of the first k-1 entries when sorted=True]

on (value, index) pairs compared by value only (ATen/native/cpu/TopKImpl.h).
Both are deterministic; below they are re-implemented from the published
libstdc++ algorithms (bits/stl_algo.h: __introselect, __move_median_to_first,
__unguarded_partition, __insertion_sort, __heap_select) so the oracle can state
the draft index the reference would pick.  `tests/test_oracle_golden.py` pins
this emulation against `torch.topk` itself on tie-heavy inputs.
"""
from __future__ import annotations

import numpy as np


def _gt(a, b):
    return a[0] > b[0]


def _move_median_to_first(v, result, a, b, c):
    if _gt(v[a], v[b]):
        if _gt(v[b], v[c]):
            v[result], v[b] = v[b], v[result]
        elif _gt(v[a], v[c]):
            v[result], v[c] = v[c], v[result]
        else:
            v[result], v[a] = v[a], v[result]
    elif _gt(v[a], v[c]):
        v[result], v[a] = v[a], v[result]
    elif _gt(v[b], v[c]):
        v[result], v[c] = v[c], v[result]
    else:
        v[result], v[b] = v[b], v[result]


def _unguarded_partition(v, first, last, pivot):
    while True:
        while _gt(v[first], v[pivot]):
            first += 1
        last -= 1
        while _gt(v[pivot], v[last]):
            last -= 1
        if not first < last:
            return first
        v[first], v[last] = v[last], v[first]
        first += 1


def _insertion_sort(v, first, last):
    if first == last:
        return
    for i in range(first + 1, last):
        val = v[i]
        if _gt(val, v[first]):
            v[first + 1:i + 1] = v[first:i]
            v[first] = val
        else:
            j = i
            while _gt(val, v[j - 1]):
                v[j] = v[j - 1]
                j -= 1
            v[j] = val


def _push_heap(v, first, hole, top, val):
    parent = (hole - 1) // 2
    while hole > top and _gt(v[first + parent], val):
        v[first + hole] = v[first + parent]
        hole = parent
        parent = (hole - 1) // 2
    v[first + hole] = val


def _adjust_heap(v, first, hole, length, val):
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if _gt(v[first + child], v[first + child - 1]):
            child -= 1
        v[first + hole] = v[first + child]
        hole = child
    if (length & 1) == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        v[first + hole] = v[first + child - 1]
        hole = child - 1
    _push_heap(v, first, hole, top, val)


def _make_heap(v, first, last):
    length = last - first
    if length < 2:
        return
    parent = (length - 2) // 2
    while True:
        _adjust_heap(v, first, parent, length, v[first + parent])
        if parent == 0:
            return
        parent -= 1


def _heap_select(v, first, middle, last):
    _make_heap(v, first, middle)
    for i in range(middle, last):
        if _gt(v[i], v[first]):
            val = v[i]
            v[i] = v[first]
            _adjust_heap(v, first, 0, middle - first, val)


def _sort_heap(v, first, last):
    while last - first > 1:
        last -= 1
        val = v[last]
        v[last] = v[first]
        _adjust_heap(v, first, 0, last - first, val)


def nth_element_desc(pairs, nth):
    """std::nth_element(first, first+nth, last, greater-by-value) on a list of (value, idx)."""
    v = pairs
    first, last = 0, len(v)
    if first == last or nth == last:
        return v
    depth = 2 * (int(last - first).bit_length() - 1)
    while last - first > 3:
        if depth == 0:
            _heap_select(v, first, nth + 1, last)
            v[first], v[nth] = v[nth], v[first]
            return v
        depth -= 1
        mid = first + (last - first) // 2
        _move_median_to_first(v, first, first + 1, mid, last - 1)
        cut = _unguarded_partition(v, first + 1, last, first)
        if cut <= nth:
            first = cut
        else:
            last = cut
    _insertion_sort(v, first, last)
    return v


def partial_sort_desc(pairs, k):
    """std::partial_sort(first, first+k, last, greater-by-value)."""
    v = pairs
    _heap_select(v, 0, k, len(v))
    _sort_heap(v, 0, k)
    return v


def topk_indices(values, k: int):
    """Indices `torch.topk(values, k, largest=True, sorted=True)` yields on the CPU backend
    for a 1-D row (only the k == 1 case is exact for k-1 > 16 sorted prefixes)."""
    vals = [(x, i) for i, x in enumerate(np.asarray(values).tolist())]
    n = len(vals)
    if k * 64 <= n:
        partial_sort_desc(vals, k)
    else:
        nth_element_desc(vals, k - 1)
        head = vals[:k - 1]
        if len(head) > 1:
            if len(head) > 16:
                raise NotImplementedError("introsort emulation for k-1 > 16 not needed by the path")
            _insertion_sort(head, 0, len(head))
            vals[:k - 1] = head
    return [i for _, i in vals[:k]]


def topk1_index(values) -> int:
    return topk_indices(values, 1)[0]
