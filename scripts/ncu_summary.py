#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into one line per captured kernel launch with the metrics the
roofline discussion needs (duration, DRAM bytes, DRAM %, tensor-pipe %, occupancy, registers, smem)."""
import csv
import re
import sys

KEYS = [
    ("dur_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_rt_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("smem_KB", "launch__shared_mem_per_block_dynamic"),
]


def to_num(v, unit, want):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.lower()
    if want.endswith("_us"):
        x *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    if want.endswith("_MB"):
        x *= {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1, "gbyte": 1e3}.get(u, 1)
    if want.endswith("_KB"):
        x *= {"byte": 1e-3, "kbyte": 1, "mbyte": 1e3}.get(u, 1)
    return round(x, 2)


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel | grid | " + " | ".join(k for k, _ in KEYS))
    for r in data:
        name = re.sub(r"\(CUtensorMap.*", "", r[idx["Kernel Name"]]).replace("void <unnamed>::", "").replace("__nv_bfloat16", "bf16")
        vals = []
        for k, m in KEYS:
            vals.append(str(to_num(r[idx[m]], units[idx[m]], k)) if m in idx else "-")
        print(f"{name.strip()} | {r[idx['Grid Size']].strip()} | " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
