"""Grids whose rows span several chunks on ONE GPU: the persistent work-list kernel (one launch for all steps) against one launch
per kernel and step.  usage: chunked_modes.py [G ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import bench

ldm = load_package()
sw, sm, b, d, dd, _ = bench.WORKLOADS["g512"]
for G in [int(a) for a in sys.argv[1:]] or [160, 200, 256, 333, 512]:
    props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
    state = bench.seeded_state(props.MomentsGrid, sm)
    steps = max(10, min(200, 40000 // G))
    out = []
    for persistent in (1, 0):
        ldm.tuning_set("reset")
        ldm.tuning_set("persistent", persistent)
        eq = ldm.Equation(props)
        eq.upload(state); eq.advance(5, 0.1); eq.synchronize()
        best = 1e30
        for _ in range(3):
            eq.upload(state); eq.advance(steps, 0.1); eq.synchronize()
            best = min(best, eq.last_elapsed_ms() / steps)
        out.append((best, eq.download().sum()))
        eq.close()
    assert out[0][1] == out[1][1]
    print(f"G {G:4d}: work-list kernel {1e3 * out[0][0]:9.2f} us per step, one launch per kernel and step {1e3 * out[1][0]:9.2f} us", flush=True)
ldm.tuning_set("reset")
