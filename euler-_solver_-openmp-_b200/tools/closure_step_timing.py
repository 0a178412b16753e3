"""Step time of the Euler loop under the three closures (ldm.h ldm_set_closure): power-1 is the reference's (persistent
kernel where the grid allows), power-2 / power-3 add the integrals over the w quadrature -- G*G*(mW-1)^2 gathers per step for
K(xi), twice that for power-2 -- with one launch per kernel and step.  Device time of ldm_advance (CUDA events on the stream);
usage: closure_step_timing.py [steps] [forced_G: only that synthetic grid]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package

ldm = load_package()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
for label, sigma, forced_G in (("14.4a", 0.15, 0), ("14.4b", 0.05, 0), ("14.4a parameters, 4 096 points", 0.15, 64),
                               ("14.4a parameters, 16 384 points", 0.15, 128)):
    if only >= 0 and forced_G != only:
        continue
    props = ldm.Properties.Generate(1, sigma, sigma, 0.4, 0.2, 0.001, forced_G)
    eq = ldm.Equation(props)
    g = np.array(props.MomentsGrid)
    X, Y = np.meshgrid(g, g, indexing="ij")
    init = np.concatenate([[200.0], (40000.0 * (1 + 0.5 * np.exp(-(X * X + Y * Y) / (2 * sigma ** 2)))).ravel()])
    samples = props.G ** 2 * (props.MaxNumForW - 1) ** 2
    for closure, name in ((1, "power-1"), (3, "power-3"), (2, "power-2")):
        eq.set_closure(closure)
        eq.upload(init); eq.advance(3, 0.1); eq.synchronize()
        eq.upload(init); eq.synchronize()
        before = eq.launches
        eq.advance(steps, 0.1); eq.synchronize()
        us = eq.last_elapsed_ms() * 1e3 / steps
        rate = "" if closure == 1 else f", {samples * (closure == 2 and 2 or 1) / us * 1e-3:7.1f} G samples/s over the w quadrature"
        print(f"{label}: G {props.G}, mW {props.MaxNumForW}, {name}: {us:9.2f} us/step, {(eq.launches - before) / steps:.1f} launches per step, "
              f"N after {steps} steps {eq.download()[0]:.12g}{rate}", flush=True)
