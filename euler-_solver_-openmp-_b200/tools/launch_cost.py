"""Fixed cost of one launch of the persistent kernels: device time of ldm_advance(n) for n = 1, 2, 5, 20, 200 steps (warm caches),
16 384-point grid.  usage: launch_cost.py [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import bench

ldm = load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "g128"
sw, sm, b, d, dd, G = bench.WORKLOADS[name]
props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
eq = ldm.Equation(props)
eq.upload(bench.seeded_state(props.MomentsGrid, sm))
eq.advance(20, 0.1); eq.synchronize()
for n in (1, 2, 5, 20, 200):
    best = 1e30
    for _ in range(5):
        eq.advance(n, 0.1); eq.synchronize()
        best = min(best, eq.last_elapsed_ms())
    print(f"{name}: advance({n:3d}): {1e3 * best:9.2f} us = {1e3 * best / n:8.2f} us per step", flush=True)
