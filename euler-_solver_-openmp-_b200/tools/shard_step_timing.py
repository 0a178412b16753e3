#!/usr/bin/env python
"""ms per sharded step on ONE rank (world = 1: the rank ships its rows to itself), one launch per step vs three.
    python tools/shard_step_timing.py [G]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

ldm = load_package()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
props = ldm.Properties.Generate(1, 0.10062893578616353, 0.08, 0.4, 0.2, 0.001, G)
g = props.MomentsGrid
state = np.empty(1 + G * G)
state[0] = 200.0
state[1:] = (200.0 ** 2 * (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.08 ** 2)))).ravel()
sh = ldm.ShardedEquation(props, 0, 1)
sh.peer_connect(lambda mine: mine.reshape(1, -1))
sh.peer_upload(state[1:], state[0])
stream = torch.cuda.current_stream()
for _ in range(3):
    sh.peer_step(0.1, stream.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(20):
    sh.peer_step(0.1, stream.cuda_stream)
e1.record(stream)
torch.cuda.synchronize()
C, N = sh.peer_download()
print(f"G={G} {os.environ.get('LDM_SHARD_ONE_LAUNCH') and 'one launch' or 'three launches'}: {e0.elapsed_time(e1) / 20:.4f} ms/step, checksum {C.sum() + N!r}")
