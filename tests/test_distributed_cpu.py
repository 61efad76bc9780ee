"""N>1 path on CPU (gloo, world_size 2): sharded packed moments all-reduced to the single-process statistics."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drloco_b200.sharding import allreduce_sum_, merge_moments, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, d, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    obs = rng.standard_normal((3, total, d)) * 2 + 0.5           # every rank draws the same global batch ...
    lo, hi = shard_range(total, rank, world)                     # ... and owns a contiguous shard of it
    mean, var, count = torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64), 1e-4
    for t in range(3):
        x = torch.from_numpy(obs[t, lo:hi])
        packed = torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([float(hi - lo)], dtype=torch.float64)])
        allreduce_sum_(packed)                                   # the one exchange of the data path
        mean, var, count = merge_moments(mean, var, count, packed[:d], packed[d:2 * d], float(packed[2 * d]))
    out[rank] = (mean.numpy(), var.numpy(), float(count))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_moments_match_single_process():
    from oracle.env_oracle import RunningMeanStd
    total, d, world = 37, 6, 2
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()        # not fork: the pytest process is multi-threaded once torch is loaded
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, total, d, out), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    obs = rng.standard_normal((3, total, d)) * 2 + 0.5
    rms = RunningMeanStd(shape=(d,))
    for t in range(3):
        rms.update(obs[t])
    for r in range(world):
        mean, var, count = out[r]
        np.testing.assert_allclose(mean, rms.mean, rtol=1e-12)
        np.testing.assert_allclose(var, rms.var, rtol=1e-10)
        assert abs(count - rms.count) < 1e-9
    # identical on every rank (bitwise): the property that keeps the replicas' normalisation in lockstep
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
