"""Host-side logic of the xi-sharded path on CPU with a real 2-rank process group (gloo): the cyclic row
ownership, the rank-major "gathered" layout and the in-place all-gather that rebuilds the replica of C each
step.  The CUDA kernels are replaced by a stand-in row update, so what is tested is exactly
the plumbing bench.py uses on GPUs: slices, permutation, collective, agreement of all ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_package


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rows_per_rank(G, world):
    return (G + world - 1) // world


def _to_gathered(C, world):
    G = C.shape[0]
    rpr = _rows_per_rank(G, world)
    out = np.zeros((world * rpr, G))
    x = np.arange(G)
    out[(x % world) * rpr + x // world] = C
    return out


def _from_gathered(Cg, G, world):
    rpr = _rows_per_rank(G, world)
    x = np.arange(G)
    return Cg[(x % world) * rpr + x // world]


def _row_update(C_full, x):
    """stand-in for one rank's kernel: new row x depends on the WHOLE old C (like the m*C integral does)"""
    return 0.5 * C_full[x] + 0.25 * C_full.mean(axis=0) + 0.01 * x


def _worker(rank, world, port, G, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rpr = _rows_per_rank(G, world)
    rng = np.random.default_rng(5)
    C0 = rng.uniform(1, 2, (G, G))
    bufs = [torch.from_numpy(_to_gathered(C0, world)), torch.zeros(world * rpr, G, dtype=torch.float64)]
    for k in range(steps):
        src, dst = bufs[k % 2], bufs[(k + 1) % 2]
        full = _from_gathered(src.numpy(), G, world)
        mine = dst[rank * rpr:(rank + 1) * rpr]
        for lx in range(rpr):
            x = rank + world * lx                      # cyclic ownership: rows rank, rank+world, ...
            if x < G:
                mine[lx] = torch.from_numpy(_row_update(full, x))
        pieces = [torch.empty(rpr * G, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(pieces, mine.reshape(-1).clone())
        dst.view(-1).copy_(torch.cat(pieces))          # rank-major: rank r's rows are contiguous at [r*rpr, (r+1)*rpr)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), _from_gathered(bufs[steps % 2].numpy(), G, world))
    dist.destroy_process_group()


@pytest.mark.parametrize("G", [12, 13])
def test_two_rank_sharded_stepping_matches_serial(tmp_path, G):
    world, steps = 2, 4
    mp.spawn(_worker, args=(world, _free_port(), G, steps, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    C = rng.uniform(1, 2, (G, G))
    for _ in range(steps):
        C = np.stack([_row_update(C, x) for x in range(G)])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"rank{r}.npy"), C)


def test_layout_helpers_of_the_package_agree(ldm):
    """ShardedEquation.to_gathered / from_gathered are the same permutation (checked without creating a handle)"""
    G, world = 13, 3
    C = np.arange(G * G, dtype=np.float64).reshape(G, G)

    class Fake:
        pass
    fake = Fake()
    fake.world, fake.G, fake.rows_per_rank = world, G, _rows_per_rank(G, world)
    got = ldm.ShardedEquation.to_gathered(fake, C)
    assert np.array_equal(got, _to_gathered(C, world))
    assert np.array_equal(ldm.ShardedEquation.from_gathered(fake, got), C)
