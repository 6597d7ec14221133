#!/usr/bin/env python
"""Aggregate an ncu source-page (SASS) CSV by CUDA source line (development tool).

usage: ncu_lines.py report.ncu-rep kernel_substring [top_n]
The SASS rows of the report are matched, in order, with ``nvdisasm --print-line-info`` output of
the same build, which carries the file/line of every instruction (-lineinfo).
"""
import csv
import os
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 45
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csrc = os.path.join(root, "document_search_engine_b200", "csrc")
cubin = os.path.join(root, "gpurun_out", "bm25f_lines.cubin")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin",
                "-o", cubin, os.path.join(csrc, "bm25f.cu")], check=True)
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
funcs = {}
cur, line = None, ("", 0)
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        funcs[cur].append(line)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
blocks, curk = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        curk = {"name": r[1], "rows": []}
        blocks.append(curk)
    elif r and r[0] == "Address":
        curk["hdr"] = r
    elif curk is not None and r:
        curk["rows"].append(r)
srcs = {}
for b in blocks:
    if kern not in b["name"]:
        continue
    cands = [k for k in funcs if kern in k]
    fn = [k for k in cands if len(funcs[k]) == len(b["rows"])]
    fn = fn[0] if fn else cands[0]
    lines = funcs[fn]
    hdr = b["hdr"]
    ie, sa = hdr.index("Instructions Executed"), hdr.index("# Samples")
    n = min(len(lines), len(b["rows"]))
    print("kernel", b["name"], "sass rows", len(b["rows"]), "disasm instrs", len(lines))
    agg = {}
    tot = tots = 0
    for i in range(n):
        r = b["rows"][i]
        e = int(r[ie] or 0)
        s_ = int(r[sa] or 0)
        a = agg.setdefault(lines[i], [0, 0, 0])
        a[0] += e
        a[1] += s_
        a[2] += 1
        tot += e
        tots += s_
    print("total warp instructions", tot, "samples", tots)
    for (f, ln), (e, s_, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        if f not in srcs:
            try:
                srcs[f] = open(os.path.join(csrc, f)).read().splitlines()
            except OSError:
                srcs[f] = []
        src = srcs[f]
        print("%5.1f%% inst %5.1f%% stall  n=%3d  %s:%-4d %s" % (100.0 * e / tot, 100.0 * s_ / max(1, tots), c, f, ln,
                                                                 src[ln - 1].strip()[:90] if 0 < ln <= len(src) else ""))
    break
