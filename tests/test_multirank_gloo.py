"""World-size-2 gloo test (CPU) of the multi-GPU host logic: sharding and the weighted
partial-accumulator all-reduce that stands in for the NCCL all-reduce of cfg 5."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
from tempestsdr_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharding_covers_everything_once():
    for n, w in [(8, 2), (30, 4), (7, 8), (1000, 8)]:
        rr = sorted(u for r in range(w) for u in parallel.shard_round_robin(n, w, r))
        assert rr == list(range(n))
        blocks = [parallel.shard_contiguous(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1


def _worker(rank, world, port, frames, alpha, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    k0, k1 = parallel.shard_contiguous(len(frames), world, rank)
    acc = np.zeros_like(frames[0])
    for k in range(k0, k1):            # the plain EMA from zero over this rank's block (what the chain does on a GPU)
        acc = orc.ema(acc, frames[k], alpha)
    t = torch.from_numpy(acc.copy())
    parallel.allreduce_partial(t, alpha, len(frames) - k1)
    # sum mode: plain frame sum, mean after the all-reduce
    s = torch.from_numpy(np.sum(np.stack(frames[k0:k1]), axis=0, dtype=np.float32).copy()) if k1 > k0 else torch.zeros_like(t)
    parallel.allreduce_partial(s, alpha, 0, sum_mode=True, total_frames=len(frames))
    if rank == 0:
        np.savez(out_path, ema=t.numpy(), mean=s.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partial_ema_allreduce(tmp_path):
    rng = np.random.default_rng(5)
    frames = [rng.random((60, 80)).astype(np.float32) for _ in range(7)]
    alpha = np.float32(0.1)
    seq = np.zeros_like(frames[0])
    for f in frames:
        seq = orc.ema(seq, f, alpha)
    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(2, _free_port(), frames, alpha, out), nprocs=2, join=True)
    got = np.load(out)
    # linear recombination of the EMA: equal to the sequential recurrence up to Float32 rounding order
    np.testing.assert_allclose(got["ema"], seq, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(got["mean"], np.mean(np.stack(frames), axis=0), rtol=2e-6, atol=1e-7)


def test_tail_weight():
    assert parallel.ema_tail_weight(0.1, 0) == 1.0
    assert abs(parallel.ema_tail_weight(0.5, 10) - 2.0 ** -10) < 1e-18
