#!/usr/bin/env python
"""Event trace of the persistent work-list kernel on the xi-sharded grid (debug build: make phase-clocks).
    python -m torch.distributed.run --nproc-per-node N ... tools/chunk_trace.py [workload] [steps]
Every rank runs the sharded leg of bench.py; ranks 0 and N-1 write gpurun_out/chunk_trace_w<N>_r<rank>.npz
(events [ctas][n][2], counts) for tools/chunk_trace_report.py."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("LDM_LIBRARY", os.path.join(ROOT, "euler-_solver_-openmp-_b200", "libldm_b200_phase.so"))
from __graft_entry__ import load_package  # noqa: E402
import bench  # noqa: E402

ldm = load_package()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
solo = int(os.environ.get("SOLO_WORLD", 0))       # one GPU plays rank SOLO_RANK of SOLO_WORLD alone (ldm_debug_shard_solo)
if solo:
    rank, world = int(os.environ.get("SOLO_RANK", 0)), solo
torch.cuda.set_device(local)
if world > 1 and not solo:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
workload = sys.argv[1] if len(sys.argv) > 1 else "g512"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sw, sm, b, d, dd, G = bench.WORKLOADS[workload]
props = ldm.Properties.Generate(1, sw, sm, b, d, dd, G)
if "ITEM_ORDER" in os.environ:
    ldm.tuning_set("item_order", int(os.environ["ITEM_ORDER"]))
sh = ldm.ShardedEquation(props, rank, world, device=local)


def all_gather_bytes(mine):
    if world == 1:
        return mine.reshape(1, -1)
    t = torch.from_numpy(mine).cuda()
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().reshape(world, -1)


def barrier():
    if world > 1 and not solo:
        dist.barrier()
    torch.cuda.synchronize()


if solo:
    ldm.library().ldm_debug_shard_solo.argtypes = [C.c_void_p]
    assert ldm.library().ldm_debug_shard_solo(sh.eq._h) == 0
else:
    sh.peer_connect(all_gather_bytes)
state0 = bench.seeded_state(props.MomentsGrid, sm)
sh.peer_upload(state0[1:], state0[0])
barrier()
stream = torch.cuda.current_stream()
sh.peer_advance(5, bench.DT, stream.cuda_stream)
torch.cuda.synchronize()
lib = ldm.library()
lib.ldm_debug_trace_reset.restype = C.c_int
lib.ldm_debug_trace_read.argtypes = [C.c_void_p, C.c_void_p]
assert lib.ldm_debug_trace_reset() == 0
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
sh.peer_advance(steps, bench.DT, stream.cuda_stream)
e1.record(stream)
torch.cuda.synchronize()
print(f"rank {rank}/{world}: {e0.elapsed_time(e1) / steps:.4f} ms per step (debug build)", flush=True)
if solo or rank in (0, world - 1):
    events = np.zeros((148, 4096, 2), dtype=np.uint64)
    counts = np.zeros(148, dtype=np.uint32)
    assert lib.ldm_debug_trace_read(events.ctypes.data, counts.ctypes.data) == 0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"chunk_trace_w{world}_r{rank}{'_solo' if solo else ''}.npz"), events=events[:, :int(counts.max())], counts=counts,
                        steps=steps, world=world, G=G)
barrier()
if world > 1 and not solo:
    dist.destroy_process_group()
