#!/usr/bin/env python
"""Compact summary of an .ncu-rep: key metrics, stall reasons, and hottest code regions per kernel."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
for r in data:
    print("==", r[idx["Kernel Name"]][:90])
    for k in KEYS:
        if k in idx:
            print(f"  {k:68s} {r[idx[k]]:>16s} {units[idx[k]]}")
    st = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            st.append((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
if len(sys.argv) > 2 and sys.argv[2] == "src":
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h = rows[1]; ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name": break
        body.append(r)
    tot = sum(int(r[isamp]) for r in body) or 1
    base = int(body[0][ia], 16)
    print(f"  code size {(int(body[-1][ia],16)-base)/1024:.1f} KB, samples {tot}")
    for r in sorted(body, key=lambda r: -int(r[isamp]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
        print(f"   {int(r[ia],16)-base:#7x} {100*int(r[isamp])/tot:5.2f}% ex={int(r[iex])//1000}k {r[isrc].strip()[:80]}")
