#!/usr/bin/env python
"""Summarise an ncu launch list (CSV with gpu__time_duration / dram bytes per launch) of a bench.py run: per-kernel
time share and DRAM traffic of one eager step; writes profiles/<tag>_launch_shares.txt and conv_dram_traffic.json."""
import collections, csv, json, re, sys
src, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h, data = rows[hi], rows[hi + 1:]
idx = {n: i for i, n in enumerate(h)}
per = {}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3}
for r in data:
    if len(r) < len(h):
        continue
    kid, name, m, u = int(r[idx["ID"]]), r[idx["Kernel Name"]], r[idx["Metric Name"]], r[idx["Metric Unit"]]
    v = float(r[idx["Metric Value"]].replace(",", ""))
    d = per.setdefault(kid, {"name": name})
    if m == "gpu__time_duration.sum":
        d["t_us"] = tscale[u] * v
    elif m == "dram__bytes_read.sum":
        d["rd"] = v * scale[u]
    elif m == "dram__bytes_write.sum":
        d["wr"] = v * scale[u]
ids = sorted(per)
starts = [k for k, i in enumerate(ids) if "image_to_nhwc4" in per[i]["name"]]
firsts = [s for j, s in enumerate(starts) if j % 7 == 0]
a, b = firsts[-2], firsts[-1]
step_all = [per[i] for i in ids[a:b]]
# the first forward of a process also builds the plans (weight packing = a few hundred tiny torch kernels): not part of a step
step = [d for d in step_all if "at::" not in d["name"]]
one_time = sum(d["t_us"] for d in step_all) - sum(d["t_us"] for d in step)
tot = sum(d["t_us"] for d in step)
def short(n):
    n = re.sub(r"void |dmvs::|\(anonymous namespace\)::|<unnamed>::|unnamed>::", "", n)
    return n.split("(")[0][:60]
fam = collections.OrderedDict()
for d in step:
    f = fam.setdefault(short(d["name"]), {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0})
    f["n"] += 1; f["t"] += d["t_us"]; f["rd"] += d.get("rd", 0); f["wr"] += d.get("wr", 0)
conv_t = sum(f["t"] for k, f in fam.items() if k.startswith("conv_"))
conv_b = sum(f["rd"] + f["wr"] for k, f in fam.items() if k.startswith("conv_"))
lines = [f"# ncu launch list of `python bench.py --steps 1 --warmup 3 --no-graph --load-tuned <table of a normal run>`",
         f"# one eager cfg3 step = launches {a}..{b} of the capture (cold-cache, serialised): {len(step)} launches, {tot / 1e3:.2f} ms",
         f"# all conv_* kernels: {100 * conv_t / tot:.1f} % of the step, DRAM traffic {conv_b / 1e9:.3f} GB per step",
         f"# (excluded: {len(step_all) - len(step)} one-time torch kernels of the plan build, {one_time / 1e3:.2f} ms)",
         "# per kernel: time, share, launches, DRAM read / written (MB)"]
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["t"]):
    lines.append(f"{f['t'] / 1e3:8.3f} ms {100 * f['t'] / tot:5.1f}% n={f['n']:4d} rd={f['rd'] / 1e6:8.1f} wr={f['wr'] / 1e6:8.1f}  {k}")
open(f"profiles/{tag}_launch_shares.txt", "w").write("\n".join(lines) + "\n")
json.dump({"conv_dram_bytes_per_step": conv_b, "conv_share_of_step_ncu": conv_t / tot, "workload": "cfg3",
           "source": f"profiles/{tag}_launch_shares.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                     "conv_* launches of one eager cfg3 step)"}, open("profiles/conv_dram_traffic.json", "w"), indent=1)
print("\n".join(lines[:24]))
