#!/usr/bin/env python
"""Turn an `ncu --set full` capture into the small JSON that bench.py reads for `roofline.traffic` and `roofline_all[*].ncu`.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2x_ncu_kernels.json

Per kernel name (averaged over the captured launches of that name): DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), duration, tensor-pipe activity, warps / issue activity, registers."""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
WANT = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes", "gpu__time_duration.sum": "duration_us",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
        "launch__registers_per_thread": "registers_per_thread", "lts__t_bytes.sum": "l2_bytes",
        "launch__grid_size": "grid", "launch__block_size": "block"}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = defaultdict(lambda: defaultdict(list))
    for r in data:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).split("::")[-1].split("<")[0].replace("void ", "").strip()
        for metric, key in WANT.items():
            if metric not in col or r[col[metric]] in ("", "n/a"):
                continue
            try:
                v = float(r[col[metric]].replace(",", ""))
            except ValueError:
                continue
            agg[name][key].append(v * UNIT.get(units[col[metric]], 1.0))
    kernels = {}
    for name, m in agg.items():
        k = {key: sum(v) / len(v) for key, v in m.items()}
        k["launches_captured"] = len(m.get("duration_us", [])) or len(next(iter(m.values())))
        k["dram_bytes_per_launch"] = k.get("dram_read_bytes", 0.0) + k.get("dram_write_bytes", 0.0)
        kernels[name] = {a: (round(b, 3) if isinstance(b, float) else b) for a, b in k.items()}
    json.dump({"source": rep, "how": "ncu --set full --clock-control none (cold-cache, serialised replays); averages per kernel name",
               "kernels": kernels}, open(out, "w"), indent=1)
    for n, k in kernels.items():
        print(f"{n:36s} {k.get('duration_us', 0):8.2f} us  dram {k['dram_bytes_per_launch'] / 1e6:7.2f} MB  tensor {k.get('tensor_pipe_active_pct_elapsed', 0):5.1f} %  "
              f"regs {int(k.get('registers_per_thread', 0))}")


if __name__ == "__main__":
    main()
