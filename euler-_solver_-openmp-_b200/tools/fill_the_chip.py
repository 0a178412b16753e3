"""The 16 384-point grid on 128 CTAs (one row per CTA, one grid barrier per step: persistent_scan_kernel, 0.86 waves) against
the same grid cut into narrower chunks and run by the work-list kernel on all 148 SMs (persistent_chunks_kernel).
usage: fill_the_chip.py [steps]   -- prints us per step and a checksum (the chunkings sum in different orders: equal to ~1e-15)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import bench

ldm = load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sw, sm, b, d, dd, G = bench.WORKLOADS["g128"]
props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
state = bench.seeded_state(props.MomentsGrid, sm)
for label, bn in (("one chunk of 128 columns, 128 CTAs + grid barrier", 0), ("2 chunks of 64 columns, work list on 148 CTAs", 64),
                  ("4 chunks of 32 columns, work list on 148 CTAs", 32)):
    ldm.tuning_set("reset")
    if bn:
        ldm.tuning_set("scan_bn", bn)
        ldm.tuning_set("persistent", 1)       # the work-list kernel (not the default on one rank)
    eq = ldm.Equation(props)
    eq.upload(state)
    eq.advance(20, 0.1); eq.synchronize()
    best = 1e30
    for _ in range(3):
        eq.upload(state)
        eq.advance(steps, 0.1); eq.synchronize()
        best = min(best, eq.last_elapsed_ms() / steps)
    print(f"{label}: {1e3 * best:8.3f} us/step, launches {eq.launches}, checksum {eq.download().sum():.12e}", flush=True)
    eq.close()
ldm.tuning_set("reset")
