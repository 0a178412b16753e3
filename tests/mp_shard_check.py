"""Two (or more) processes, one GPU each: the xi-sharded step with the exchange fused in (ldm_shard_peer_advance) against
the same steps on one rank, ELEMENT BY ELEMENT.  Run under torchrun by tests/test_gpu_parity.py and by hand:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tests/mp_shard_check.py
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--G", type=int, default=256)
    ap.add_argument("--steps", type=int, default=12)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    ldm = load_package()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    props = ldm.Properties.Generate(1, 0.10062893578616353, 0.08, 0.4, 0.2, 0.001, args.G)
    g = props.MomentsGrid
    state = np.full(1 + args.G * args.G, 200.0 ** 2)
    state[0] = 200.0
    state[1:] *= (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.08 ** 2)) + 0.05 * np.sin(5 * g[:, None] + 2 * g[None, :])).ravel()
    sh = ldm.ShardedEquation(props, rank, world, device=local)

    def all_gather_bytes(mine):
        t = torch.from_numpy(mine).cuda()
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().reshape(world, -1)

    sh.peer_connect(all_gather_bytes)
    sh.peer_upload(state[1:], state[0])
    dist.barrier()
    stream = torch.cuda.current_stream().cuda_stream
    sh.peer_advance(args.steps // 2, 0.1, stream)
    sh.peer_advance(args.steps - args.steps // 2, 0.1, stream)
    got_C, got_N = sh.peer_download()
    dist.barrier()
    single = ldm.Equation(props, device=local)
    single.upload(state)
    single.advance(args.steps, 0.1)
    want = single.download()
    diff = float(np.max(np.abs(got_C.ravel() - want[1:])))
    print(f"rank {rank}/{world}: G={args.G} steps={args.steps} max_abs_diff {max(diff, abs(got_N - want[0]))}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if diff != 0.0 or got_N != want[0]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
