#!/usr/bin/env python
"""Summarise an ncu report (CPU side): headline metrics + the top stall hot spots of the source page.
    python scripts/ncu_hot.py gpurun_out/prof_x.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__grid_size", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} = {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {c: sum(int(r[ix[c]] or 0) for r in data) for c in stall}
tot = sum(agg.values()) or 1
print("stalls:", ", ".join(f"{c[6:]} {100 * v / tot:.0f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:topn]:
    st = sorted(((c, int(r[ix[c]] or 0)) for c in stall), key=lambda kv: -kv[1])[:2]
    print(r[ix["Address"]][-5:], r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]][:72].ljust(72), [(c[6:], v) for c, v in st])
