#!/usr/bin/env python
"""Phase timeline of the prefix-sum kernels (debug build: make -C euler-_solver_-openmp-_b200 phase-clocks).
Prints, averaged over the CTAs of the last step / launch, the SM-clock cycles spent between the phase marks.
    python tools/scan_clocks.py [G] [sigma_w sigma_m]"""
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
sw, sm = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0.15, 0.15)
props = ldm.Properties.Generate(1, sw, sm, 0.4, 0.2, 0.001, G)
if os.environ.get("LDM_CLOCKS_PER_STEP_LAUNCHES"):      # the per-step kernel instead of the persistent ones
    ldm.tuning_set("persistent", 0)
eq = ldm.Equation(props)
st = eq.GenerateEquilibriumValuesForNAndC(200)
eq.upload(st)
eq.advance(20, 0.1)
eq.synchronize()
eq.advance(200, 0.1)
eq.synchronize()
print(f"G={G} A={(props.MaxNumForM - 1) // 2}: {eq.last_elapsed_ms() / 200 * 1e3:.2f} us per step over 200 steps (this debug build)")
rows = 148 * 8
buf = np.zeros((rows, 12), dtype=np.int64)
lib = ldm.library()
# one timeline per translation unit: the persistent kernels' (ldm_persist.cu) where one ran, else the per-step kernel's
lib.ldm_debug_persist_clocks.argtypes = [C.c_void_p, C.c_int]
lib.ldm_debug_scan_clocks.argtypes = [C.c_void_p, C.c_int]
assert lib.ldm_debug_persist_clocks(buf.ctypes.data, rows) == 0
if not np.any(buf[:, 1] > 0):
    assert lib.ldm_debug_scan_clocks(buf.ctypes.data, rows) == 0
names = ["tables (first step only)", "E + running sum + select", "(unused)", "y side", "write-out", "WC rows",
         "(to barrier 1)", "barrier 1", "closing phase", "barrier 2"]
body = buf[buf[:, 1] > 0]
print(f"{len(body)} CTAs ran the row body; body total {np.mean(body[:, 5] - body[:, 0]):.0f} cycles (max {np.max(body[:, 5] - body[:, 0])})")
for i, n in enumerate(names[:5]):
    col = body[:, i + 1] - body[:, i] if i != 1 else body[:, 3] - body[:, 1]
    print(f"  {n:28s} mean {col.mean():9.0f}  max {col.max():9.0f} cycles")
pers = buf[(buf[:, 9] > 0) & (buf[:, 1] > 0)]
if len(pers):
    print(f"persistent kernel, {len(pers)} row CTAs, last step:")
    for n, (i, j) in {"ring correction of the row": (5, 6), "closing of the row + WC row": (6, 8), "grid barrier": (8, 9)}.items():
        col = pers[:, j] - pers[:, i]
        print(f"  {n:28s} mean {col.mean():9.0f}  max {col.max():9.0f} cycles")
    col = pers[:, 11] - pers[:, 10]
    print(f"  {'ring share (step start)':28s} mean {col.mean():9.0f}  max {col.max():9.0f} cycles")
    print(f"  step (ring share + mark 0 -> 9) mean {np.mean(pers[:, 9] - pers[:, 0] + col):9.0f} cycles")
if len(pers) and os.environ.get("LDM_CLOCKS_PER_ROW"):
    # per row CTA: ring share, E, y side, write-out + closing, barrier wait (cycles of the last step)
    print("row   ring     E      y    close  barrier   SM  share: CSR rows, entries")
    for x in range(len(pers)):
        r = pers[x]
        print(f"{x:3d} {r[11] - r[10]:6d} {r[3] - r[1]:6d} {r[4] - r[3]:6d} {r[8] - r[4]:6d} {r[9] - r[8]:6d}  {r[7]:4d}  {r[2] >> 32:3d} {r[2] & 0xffffffff:6d}")
