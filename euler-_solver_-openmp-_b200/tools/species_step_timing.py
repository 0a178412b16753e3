"""Step time of the multi-species Euler loop on the reference's 2- and 3-species example inputs: ldm_advance (two steps per
replayed CUDA graph) against record = all of ldm_solve (one launch sequence per step, history copied out).  Wall clock
around synchronised calls; usage: species_step_timing.py [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package

ldm = load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for s, sw in ((2, [[0.05, 0.06], [0.06, 0.05]]), (3, [[0.05, 0.06, 0.07], [0.06, 0.05, 0.04], [0.07, 0.04, 0.05]])):
    props = ldm.Properties.Generate(s, sw, [0.05] * s, [0.4, 0.401, 0.4001][:s], [0.2] * s, np.full((s, s), 0.001))
    eq = ldm.Equation(props)
    init = eq.GenerateEquilibriumValuesForNAndC(200)
    sums = []
    for label, graph in (("replayed graph", 1), ("ordinary launches", 0)):
        ldm.tuning_set("multi_graph", graph)
        eq.upload(init); eq.advance(50, 0.1); eq.synchronize()
        eq.upload(init); eq.synchronize()
        before = eq.launches
        t0 = time.perf_counter(); eq.advance(steps, 0.1); eq.synchronize(); wall = time.perf_counter() - t0
        sums.append(eq.download().sum())
        print(f"{s} species, G {props.G}, {label}: {1e6 * wall / steps:7.2f} us/step wall, {eq.last_elapsed_ms() * 1e3 / steps:7.2f} us/step device, "
              f"{(eq.launches - before) / steps:.1f} launches per step", flush=True)
    assert sums[0] == sums[1]
    ldm.tuning_set("reset")
