#!/usr/bin/env python
"""What ONE rank of a W-rank xi-sharded run computes per step, without any exchange (ldm_shard_step on one GPU):
the part of the sharded step that scales.    python tools/shard_rank_timing.py [G] [W ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

ldm = load_package()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
worlds = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
props = ldm.Properties.Generate(1, 0.10062893578616353, 0.08, 0.4, 0.2, 0.001, G)
g = props.MomentsGrid
C0 = (200.0 ** 2 * (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.08 ** 2))))
stream = torch.cuda.current_stream()
for W in worlds:
    sh = ldm.ShardedEquation(props, 0, W)
    rpr = sh.rows_per_rank
    src = torch.from_numpy(sh.to_gathered(C0.ravel())).cuda()
    dst = torch.zeros(rpr, G, dtype=torch.float64, device="cuda")
    sh.set_N(200.0)
    for _ in range(3):
        sh.step(src.data_ptr(), dst.data_ptr(), 0.1, stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        sh.step(src.data_ptr(), dst.data_ptr(), 0.1, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"G={G} world={W}: rank 0 computes its {rpr} rows in {e0.elapsed_time(e1) / 20:.4f} ms per step")
