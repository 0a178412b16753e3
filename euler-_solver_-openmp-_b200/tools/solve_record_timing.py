"""Where the time of EulerSolver::Solve's twin goes by record mode (16 384-point grid, 200 steps, pinned host buffers):
wall clock of the call against the device time of its stepping loop.  usage: solve_record_timing.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import bench

ldm = load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sw, sm, b, d, dd, G = bench.WORKLOADS["g128"]
props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
eq = ldm.Equation(props)
init = torch.from_numpy(bench.seeded_state(props.MomentsGrid, sm)).pin_memory().numpy()
width = 1 + G * G
vals = torch.empty((steps + 1) * width, dtype=torch.float64).pin_memory().numpy()
ders = torch.empty((steps + 1) * width, dtype=torch.float64).pin_memory().numpy()
solver = ldm.EulerSolver()
for name, record in (("last", ldm.RECORD_LAST), ("observables", ldm.RECORD_OBSERVABLES), ("values", ldm.RECORD_VALUES), ("all", ldm.RECORD_ALL)):
    kw = dict(values_out=vals, derivs_out=ders) if record >= ldm.RECORD_VALUES else {}
    solver.Solve(eq, ldm.SquaredN, (20 + 0.5) * 0.1, 0.1, init, record=record, **kw)
    runs = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.Solve(eq, ldm.SquaredN, (steps + 0.5) * 0.1, 0.1, init, record=record, **kw)
        wall = time.perf_counter() - t0
        runs.append((wall, eq.last_elapsed_ms()))
    print(f"record {name:12s}: " + "; ".join(f"wall {1e6 * w / steps:6.2f} device {1e3 * dev / steps:6.2f}" for w, dev in runs) + "  us/step (three calls in a row)", flush=True)
