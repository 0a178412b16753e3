#!/usr/bin/env python
"""Top CUDA source lines of one kernel from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: source_hotspots.py file.csv [n_lines]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if r and r[0] == "Line No" and len(r) > 10)
col = {h: i for i, h in enumerate(hdr)}
samp, inst = col["# Samples"], col["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
fname = "?"
for r in rows:
    if r and r[0] in ("File Name", "File Path") and len(r) > 1:
        fname = r[1].split("/")[-1]
    if len(r) <= samp or r[2] != "-" or not r[0].isdigit():
        continue
    try:
        s, n = float(r[samp]), float(r[inst])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    e = agg.setdefault(key, [0.0, 0.0, r[1].strip(), {}])
    e[0] += s; e[1] += n
    for h in stalls:
        try:
            e[3][h] = e[3].get(h, 0) + float(r[col[h]])
        except ValueError:
            pass
tot_s = sum(e[0] for e in agg.values()) or 1
tot_i = sum(e[1] for e in agg.values()) or 1
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
for (f, line), e in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(e[3].items(), key=lambda kv: -kv[1])[:3]
    st = " ".join(f"{k[6:]}={v:.0f}" for k, v in st if v > 0)
    print(f"{100 * e[0] / tot_s:5.1f}% smp {100 * e[1] / tot_i:5.1f}% inst  {f}:{line:<4} {e[2][:80]}\n{'':28}{st}")
