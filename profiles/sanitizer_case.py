"""Small end-to-end case for compute-sanitizer: both algorithms, a multi-chunk tiling, the batch launch, the sharded layout."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

ldm = load_package()
props = ldm.Properties.Generate(1, 0.05, 0.05, 0.4, 0.2, 0.001)
for algo in (ldm.ALGO_DIRECT, ldm.ALGO_FACTORED):
    eq = ldm.Equation(props, algorithm=algo)
    res = ldm.EulerSolver().Solve(eq, ldm.SquaredN, 0.55, 0.1, eq.GenerateEquilibriumValuesForNAndC(200), record=ldm.RECORD_ALL)
    print("algo", algo, res.Values[-1][:3])
os.environ["LDM_FACTORED_BN"] = "64"
os.environ["LDM_FACTORED_BM"] = "40"
big = ldm.Properties.Generate(1, 0.12, 0.09, 0.4, 0.2, 0.001, 96)
eq = ldm.Equation(big)
eq.upload(eq.GenerateEquilibriumValuesForNAndC(200))
eq.advance(2, 0.1)
print("tiled", eq.download()[:3])
eqs = [ldm.Equation(ldm.Properties.Generate(1, s, 0.06, 0.4, 0.2, 0.001)) for s in (0.03, 0.05, 0.09)]
outs = ldm.solve_batch(eqs, [e.GenerateEquilibriumValuesForNAndC(200) for e in eqs], 0.1, 0.35)
print("batch", [o[0] for o in outs])
