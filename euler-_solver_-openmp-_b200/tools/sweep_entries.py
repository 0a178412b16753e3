import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
ldm = load_package()
sigmas = np.linspace(1e-8, 0.2, 160); sigmas[-1] = 0.2
for i in (0, 5, 11, 12, 80, 159):
    p = ldm.Properties.Generate(1, float(sigmas[i]), 0.08, 0.4, 0.2, 0.001)
    eq = ldm.Equation(p)
    init = eq.GenerateEquilibriumValuesForNAndC(200)
    eq.upload(init); eq.advance(100, 0.1); eq.synchronize()
    eq.upload(init); eq.advance(1000, 0.1); eq.synchronize()
    a = eq.last_elapsed_ms()
    ldm.solve_batch([eq], [init], 0.1, 10.05)
    ldm.solve_batch([eq], [init], 0.1, 100.05)
    b = eq.last_elapsed_ms()
    print(f"entry {i}: G {p.G} K {p.K} mM {p.MaxNumForM} mW {p.MaxNumForW}: advance {a:7.2f} us/step, batch of one {b:7.2f} us/step", flush=True)
