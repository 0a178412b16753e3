"""The roofline denominators measured live on this device, five times in a row: FP64 (DFMA chain), load/store data-pipe
wavefronts, L2 -> SM read bandwidth."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from __graft_entry__ import load_package

ldm = load_package()
for _ in range(5):
    wf, l2 = ldm.measure_memory_pipes(0)
    print(f"FP64 {ldm.measure_fp64_peak(0):6.2f} TFLOP/s, load/store pipe {wf / 1e9:7.2f} G wavefronts/s, L2 -> SM {l2:8.1f} GB/s", flush=True)
