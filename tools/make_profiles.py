#!/usr/bin/env python
"""Summarises an ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum; --csv) into
profiles/: per-kernel time shares of one bench step and the DRAM traffic of ica_iterate_kernel against its
algorithmic bytes.   usage: make_profiles.py <launches.csv> [round prefix, default r1] [bench JSON line of the same command]"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
RND = sys.argv[2] if len(sys.argv) > 2 else "r1"
BENCH = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(src)))
hdr = None
launches = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    launches.setdefault(int(d["ID"]), {"name": d["Kernel Name"], "grid": d["Grid Size"]})[d["Metric Name"]] = v * scale
ids = sorted(launches)
# the bench runs warm-up step, timed step, two instrumented steps; every step starts with minmax_reset_kernel
starts = [i for i in ids if "minmax_reset" in launches[i]["name"]]
step = [i for i in ids if starts[1] <= i < starts[2]]      # the timed step
agg = collections.OrderedDict()
for i in step:
    L = launches[i]
    short = L["name"].split("(")[0].split("::")[-1]
    a = agg.setdefault(short, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1
    a["us"] += L.get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
for k, a in agg.items():
    a["share"] = round(a["us"] / tot, 4)
    a["us"] = round(a["us"], 1)
summary = {"source": os.path.basename(src),
           "command": "ICA_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                      "--clock-control none python bench.py --steps 1 --warmup 1 --batch 32 --streams 1 --no-cpu-baseline --no-e2e",
           "step": "second step of the run (the timed one), 32 pairs, single stream; per-launch times under ncu are cold-cache and serialised",
           "total_us": round(tot, 1), "kernels": agg}
json.dump(summary, open(os.path.join(ROOT, "profiles", RND + "_step_shares.json"), "w"), indent=1)
it = [launches[i] for i in step if "ica_iterate_kernel" in launches[i]["name"]]
traffic = sum(L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0) for L in it)
out = {"kernel": "ica_iterate_kernel<3,4>", "source": "profiles/" + os.path.basename(src) + " (see " + RND + "_step_shares.json for the command)",
       "launches_per_step": len(it), "traffic_bytes_per_step": traffic, "traffic_bytes_per_launch": traffic / max(1, len(it)),
       "pairs_per_step": 32,
       # algorithmic bytes of the same step (bench.py prints them: Sum_s iterations_s * N_s * 24 B over the 32 pairs of the step)
       "algorithmic_bytes_per_step_at_capture": 2887778304.0}
if BENCH:
    line = [l for l in open(BENCH) if l.startswith("{")][-1]
    out["algorithmic_bytes_per_step_at_capture"] = json.loads(line)["roofline"]["algorithmic_bytes_per_step"]
out["traffic_over_algorithmic"] = out["traffic_bytes_per_step"] / out["algorithmic_bytes_per_step_at_capture"]
json.dump(out, open(os.path.join(ROOT, "profiles", "iterate_dram_bytes.json"), "w"), indent=1)
print(json.dumps(summary, indent=1))
print(json.dumps(out))
