"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): the prefix-sum path in its three launch regimes (one-chunk
persistent kernel with its shared-memory copies, work-list kernel on chunked rows, per-step launches), the dense contraction
and the direct sum, the persistent batch, several species (replayed graph), history straight into pinned host buffers.  (Ranks
side by side on one device are left out: the tool serialises kernels, and a rank would wait for a peer that cannot start.)
    compute-sanitizer --tool memcheck python profiles/sanitizer_case.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

ldm = load_package()
props = ldm.Properties.Generate(1, 0.05, 0.05, 0.4, 0.2, 0.001)
for algo in (ldm.ALGO_DIRECT, ldm.ALGO_FACTORED, ldm.ALGO_PREFIX):
    eq = ldm.Equation(props, algorithm=algo)
    res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 0.55, 0.1, eq.GenerateEquilibriumValuesForNAndC(200), record=ldm.RECORD_ALL)
    print("algo", algo, res.Values[-1][:3])
mid = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, 96)
eq = ldm.Equation(mid)                                    # 512-thread one-chunk persistent kernel: ring share, cached rows
init = eq.GenerateEquilibriumValuesForNAndC(200)
width = 1 + mid.G ** 2
values = torch.empty(4 * width, dtype=torch.float64).pin_memory().numpy()
derivs = torch.empty(4 * width, dtype=torch.float64).pin_memory().numpy()
res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 0.35, 0.1, init, record=ldm.RECORD_ALL, values_out=values, derivs_out=derivs)
print("pinned history", res.Values[-1][:2], res.Derivatives[-1][:2])
ldm.tuning_set("scan_bn", 32)
chunked = ldm.Equation(mid)                               # three chunks per row: the work-list kernel
ldm.tuning_set("persistent", 0)
stepwise = ldm.Equation(mid)                              # the same, one launch per kernel and step
ldm.tuning_set("reset")
for e in (chunked, stepwise):
    e.upload(init); e.advance(3, 0.1)
print("chunked", np.array_equal(chunked.download(), stepwise.download()))
eqs = [ldm.Equation(ldm.Properties.Generate(1, s, 0.06, 0.4, 0.2, 0.001)) for s in (0.03, 0.05, 0.09)]
outs = ldm.solve_batch(eqs, [e.GenerateEquilibriumValuesForNAndC(200) for e in eqs], 0.1, 0.35)
print("batch", [o[0] for o in outs])
two = ldm.Properties.Generate(2, [[0.05, 0.06], [0.06, 0.05]], [0.05] * 2, [0.4, 0.401], [0.2] * 2, np.full((2, 2), 0.001))
multi = ldm.Equation(two)
multi.upload(multi.GenerateEquilibriumValuesForNAndC(200)); multi.advance(12, 0.1)
print("species", multi.download()[:2])
