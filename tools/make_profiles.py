#!/usr/bin/env python
"""Turn the scratch outputs of a measurement call (gpurun_out/) into the tracked files under profiles/.

Expects: gpurun_out/bench_n1.json, gpurun_out/launches_r01.csv (ncu launch list of bench.py, own kernels),
gpurun_out/prof_r01_bench.ncu-rep (ncu --set full of k_score_stream / k_score_isect in bench.py)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "launches_r01.csv"), os.path.join(P, "r01_launches_bench_n1.csv"))
shutil.copy(os.path.join(G, "bench_n1.json"), os.path.join(P, "r01_bench_n1.json"))
rows = [r for r in csv.reader(open(os.path.join(G, "launches_r01.csv"))) if r and not r[0].startswith("==")]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
L = collections.OrderedDict()
for r in rows[1:]:
    d = L.setdefault(r[iid], {"k": re.search(r"k_[a-z_]+", r[ik]).group(0)})
    d[r[im]] = float(r[iv].replace(",", ""))
names = []
for d in L.values():
    if d["k"] not in names:
        names.append(d["k"])


def avg(k, m):
    ds = [d for d in L.values() if d["k"] == k]
    return sum(d[m] for d in ds) / len(ds)


per = {k: {"ms": avg(k, "gpu__time_duration.sum") / 1e6,
           "dram_bytes": avg(k, "dram__bytes_read.sum") + avg(k, "dram__bytes_write.sum"),
           "launches_profiled": len([d for d in L.values() if d["k"] == k])} for k in names}
b = json.load(open(os.path.join(G, "bench_n1.json")))
serial = sum(v["ms"] * (2 if k == "k_decode_keys" else 1) for k, v in per.items())
scoring = sum(v["ms"] for k, v in per.items() if k.startswith("k_score"))
pk = b["roofline"]["whole_step"]["postings_by_kernel"]
out = {"command": "python bench.py --steps 3 --warmup 3 --no-cpu --check 0 under ncu --metrics gpu__time_duration.sum,"
                  "dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_(score|merge|decode|tile) -c 48 "
                  "(launches are serialised under ncu; k_score_stream and k_score_isect overlap in a normal run)",
       "per_step": per, "serialised_step_ms": serial, "scoring_share_of_serialised_step": scoring / serial,
       "algorithmic_bytes": {"k_score_stream": 9 * pk["k_score_stream"], "k_score_isect": 9 * pk["k_score_isect"],
                             "k_score_team": 9 * pk["k_score_team"]},
       "bench_cuda_events": {"step_ms": b["ms_per_step"], "scoring_ms": b["roofline"]["whole_step"]["ms"],
                             "k_score_stream_ms": b["roofline"]["kernel_ms_per_step"], "merge_ms": b["kernel_ms"]["merge"]}}
json.dump(out, open(os.path.join(P, "r01_traffic.json"), "w"), indent=1)
rep = os.path.join(G, "prof_r01_bench.ncu-rep")
with open(os.path.join(P, "r01_ncu_full_bench_n1.txt"), "w") as f:
    for k in ("stream", "isect"):
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, k], capture_output=True, text=True).stdout)
    for k in ("k_score_stream", "k_score_isect"):
        f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, k, "25"], capture_output=True, text=True).stdout)
print(json.dumps(out["per_step"], indent=1))
print("serialised", serial, "share", scoring / serial)
