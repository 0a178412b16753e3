#!/usr/bin/env python
"""How many loads of C the E phase of the prefix-sum path issues per (row x, column) and how many of them touch distinct rows
of C -- host tables only, no GPU.  usage: python profiles/e_phase_reuse.py > profiles/e_phase_reuse.txt

  issued      4 taps per quadrature node alpha (what scan_body does today; zero-weight taps beyond the grid included)
  block U     register reuse between the U consecutive alphas one thread takes per super-chunk (a tap pair that equals the
              pair before costs nothing, one that has moved by one row costs one load)
  sliding     the same with one thread walking ALL alphas of a side in order
  distinct    rows of C that the row's taps touch at all
"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ldm = importlib.import_module("euler-_solver_-openmp-_b200")
import bench


def walk(up, A, block):
    loads = 0
    for a0 in range(0, A, block):
        prev = [None, None]
        for a in range(a0, min(A, a0 + block)):
            for side, o in ((0, A - 1 - a), (1, A + a)):
                u = up[o]
                if u == 0:
                    prev[side] = None
                    loads += 2
                    continue
                if prev[side] is None or abs(u - prev[side]) > 1:
                    loads += 2
                elif u != prev[side]:
                    loads += 1
                prev[side] = u
    return loads


for name in ("g128", "g256", "g512"):
    sw, sm, b, d, dd, G = bench.WORKLOADS[name]
    p = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
    g, kg, m = p.MomentsGrid, p.KernelsGrid, p.MaxNumForM
    offs = np.array([kg[i + 1] if i >= 0 else -kg[-i - 1] for i in range(1 - m, m - 1, 2)])
    A = len(offs) // 2
    rows = {"issued": 0, "distinct": 0, "sliding": 0}
    blocks = {U: 0 for U in (4, 5, 8, 16)}
    for x in range(G):
        s = np.abs(g[x] + offs)
        up = np.maximum(np.searchsorted(g, s, side="left"), 1) * (s <= p.CutoffDistance)
        rows["issued"] += 4 * A
        rows["distinct"] += len({int(u) for u in up if u} | {int(u) - 1 for u in up if u})
        rows["sliding"] += walk(up, A, A)
        for U in blocks:
            blocks[U] += walk(up, A, U)
    print(f"{name}: G {G}, A {A}: loads per column and row: issued {rows['issued'] / G:.0f}, "
          + ", ".join(f"block U={U} {v / G:.0f}" for U, v in blocks.items())
          + f", sliding {rows['sliding'] / G:.0f}, distinct rows {rows['distinct'] / G:.0f}")
