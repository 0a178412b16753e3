"""CTA size of the small problems (sweep entries, batch members): 256 threads (default) against 128.
usage: small_cta_experiment.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package
import bench

ldm = load_package()
for nt in (0, 128):
    ldm.tuning_set("reset")
    if nt:
        ldm.tuning_set("scan_nt", nt)
    # the 14.6 a sweep
    sigmas = np.linspace(1e-8, 0.2, 160); sigmas[-1] = 0.2
    eqs = [ldm.Equation(ldm.Properties.Generate(1, float(sw), 0.08, 0.4, 0.2, 0.001)) for sw in sigmas]
    inits = [e.GenerateEquilibriumValuesForNAndC(200) for e in eqs]
    outs = ldm.solve_batch(eqs, inits, 0.1, 200)
    print(f"nt {nt or 'default'}: sweep stepping {eqs[0].last_elapsed_ms():8.2f} ms, N_end {[float(outs[i][0]) for i in (0, 12, 80, 159)]}", flush=True)
    for e in eqs:
        e.close()
    # 128 sets of 4096 points
    sets = bench.lattice_parameter_sets(128)
    eqs = [ldm.Equation(ldm.Properties.Generate(1, sigma, sigma, b, d, 0.001, 64)) for (sigma, b, d) in sets]
    if eqs:
        inits = [bench.seeded_state(e.props.MomentsGrid, s[0]) for e, s in zip(eqs, sets)]
        outs = ldm.solve_batch(eqs, inits, 0.1, 5.05)
        outs = ldm.solve_batch(eqs, inits, 0.1, 5.05)
        print(f"nt {nt or 'default'}: batch 128 x 4096: {eqs[0].last_elapsed_ms() / 50 * 1e3:8.2f} us/step, checksum {sum(o.sum() for o in outs):.10e}", flush=True)
        for e in eqs:
            e.close()
ldm.tuning_set("reset")
