#!/usr/bin/env python
"""Where a step of the persistent work-list kernel goes: reads gpurun_out/chunk_trace_*.npz (tools/chunk_trace.py).
    python tools/chunk_trace_report.py FILE.npz [first_step last_step]"""
import sys

import numpy as np

d = np.load(sys.argv[1])
ev, counts, steps = d["events"], d["counts"], int(d["steps"])
k0, k1 = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5, steps - 5)
t0 = min(int(ev[c, 0, 0]) for c in range(len(counts)) if counts[c])
spans = {"wait (item start -> E phase)": [], "item body": [], "closing: wait for the state": [], "closing: work": [], "ring share": []}
step_first, step_last_close = {}, {}
busy = 0.0
per_cta_items = []
for c in range(len(counts)):
    n = int(counts[c])
    if n == 0:
        continue
    t = ev[c, :n, 0].astype(np.int64) - t0
    w = ev[c, :n, 1]
    typ, k = (w >> np.uint64(48)).astype(int), ((w >> np.uint64(32)) & np.uint64(0xffff)).astype(int)
    items = 0
    last = {}
    for i in range(n):
        ty, kk, ti = typ[i], k[i], t[i]
        if ty == 1:
            last[1] = ti
        elif ty == 2:
            if k0 <= kk < k1:
                spans["wait (item start -> E phase)"].append(ti - last.get(1, ti))
            last[2] = ti
            step_first[kk] = min(step_first.get(kk, ti), ti)
        elif ty == 3:
            if k0 <= kk < k1:
                spans["item body"].append(ti - last[2]); items += 1
        elif ty == 4:
            last[4] = ti
        elif ty == 5:
            if k0 <= kk < k1:
                spans["closing: wait for the state"].append(ti - last[4])
            last[5] = ti
        elif ty == 6:
            if k0 <= kk < k1:
                spans["closing: work"].append(ti - last[5])
            step_last_close[kk] = max(step_last_close.get(kk, ti), ti)
        elif ty == 7:
            pass
    per_cta_items.append(items)
ks = sorted(step_last_close)
period = np.diff([step_last_close[k] for k in ks if k0 <= k < k1])
n_steps = k1 - k0
ctas = len(per_cta_items)
print(f"{sys.argv[1]}: {ctas} CTAs, steps {k0}..{k1 - 1}; step period (last closing to last closing) {period.mean() / 1e3:.1f} us "
      f"(min {period.min() / 1e3:.1f}, max {period.max() / 1e3:.1f})")
tot = period.mean() * ctas
for name, v in spans.items():
    v = np.asarray(v, dtype=np.float64)
    if len(v) == 0:
        continue
    print(f"  {name:32s} n/step {len(v) / n_steps:7.1f}  mean {v.mean() / 1e3:7.2f} us  max {v.max() / 1e3:7.2f}  "
          f"share of CTA time {100 * v.sum() / n_steps / tot:5.1f} %")
print(f"  items per CTA and step: mean {np.mean(per_cta_items) / n_steps:.2f}  min {np.min(per_cta_items) / n_steps:.2f}  max {np.max(per_cta_items) / n_steps:.2f}")
lead = [step_first[k + 1] - step_last_close[k] for k in ks if k + 1 in step_first and k0 <= k < k1]
print(f"  first E phase of step k+1 starts {-np.mean(lead) / 1e3:.1f} us before the last closing of step k ends")
