#!/usr/bin/env python
"""Phase timeline of factored_kernel (debug build: make -C euler-_solver_-openmp-_b200 phase-clocks).
Prints, averaged over the CTAs of one launch, the SM-clock cycles spent between the phase marks."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("LDM_LIBRARY", os.path.join(ROOT, "euler-_solver_-openmp-_b200", "libldm_b200_phase.so"))
from __graft_entry__ import load_package  # noqa: E402

ldm = load_package()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 128
props = ldm.Properties.Generate(1, 0.15, 0.15, 0.4, 0.2, 0.001, G)
eq = ldm.Equation(props)
st = eq.GenerateEquilibriumValuesForNAndC(200)
eq.upload(st)
eq.advance(20, 0.1)
eq.synchronize()
rows = min(148 * 4, G * 4)
buf = np.zeros((rows, 12), dtype=np.int64)
lib = ldm.library()
lib.ldm_debug_phase_clocks.argtypes = [C.c_void_p, C.c_int]
assert lib.ldm_debug_phase_clocks(buf.ctypes.data, rows) == 0
names = ["prologue (tap tables)", "first E/Q chunk", "K loop", "stage H tile", "y-side contraction", "write-out + WC rows"]
buf = buf[buf[:, 0] > 0]
d = np.diff(buf[:, :7], axis=1)
print("K_x per row:", "n/a"); print(f"{len(buf)} CTAs; total {np.mean(buf[:, 6] - buf[:, 0]):.0f} cycles (max {np.max(buf[:, 6] - buf[:, 0])})")
for n, col in zip(names, d.T):
    print(f"  {n:28s} mean {col.mean():9.0f}  max {col.max():9.0f} cycles")
print(f"  K loop: consumer waited {buf[:, 7].mean():.0f} cycles for full buffers, producer waited {buf[:, 8].mean():.0f} cycles for empty buffers")
