#!/usr/bin/env python
"""Multi-GPU checks of the statistics exchange and the data-parallel learner (run under torchrun on >= 2 GPUs):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
         tools/multi_gpu_check.py

1. the peer-mailbox exchange (csrc/vecnorm.cu::vecnorm_step_kernel) comes up and every rank ends with bit-identical
   running statistics; they equal the NCCL-all-reduce path's to rounding, and a single process that steps all the
   shards itself (world-size-1 semantics of SB3's VecNormalize over the whole batch) to rounding;
2. stats_sync_every > 1 keeps the ranks in lockstep too;
3. the data-parallel PPO replicas stay bit-identical over updates on real rollouts.
Prints one JSON line on rank 0; exit code 0 iff everything held.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from drloco_b200.vec_env import B200MimicVecEnv, B200VecNormalize  # noqa: E402


def run(exchange, n, steps, rank, local, every=1):
    env = B200MimicVecEnv("StraightMimicWalker", num_envs=n, device=f"cuda:{local}", seed=5, env_id_offset=rank * n)
    vn = B200VecNormalize(env, exchange=exchange, stats_sync_every=every)
    g = torch.Generator(device=env.device)
    g.manual_seed(100 + rank)
    vn.reset_tensor()
    for k in range(steps):
        vn.step_tensor(torch.rand(n, 8, device=env.device, generator=g) * 2 - 1)
    torch.cuda.synchronize()
    rms = vn._rms[vn._cur].clone()
    mode = vn.exchange
    vn.close()
    return rms, mode


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=90))
    n, steps = 1024, 25
    out = {"world": world}
    ok = True

    def gathered(t):
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        return lst

    rms_peer, mode = run("auto", n, steps, rank, local)
    out["exchange"] = mode
    ident = all(torch.equal(x, rms_peer) for x in gathered(rms_peer))
    out["peer_identical_across_ranks"] = ident
    ok &= ident and mode == "peer"
    rms_nccl, mode2 = run("nccl", n, steps, rank, local)
    d = float(((rms_peer - rms_nccl).abs() / rms_nccl.abs().clamp(min=1e-12)).max())
    out["peer_vs_nccl_max_rel"] = d
    ok &= d < 1e-9 and mode2 == "nccl"
    rms_k, _ = run("auto", n, steps - 1, rank, local, every=4)          # 24 steps: merged at 4, 8, ..., 24
    identk = all(torch.equal(x, rms_k) for x in gathered(rms_k))
    out["sync_every_4_identical_across_ranks"] = identk
    ok &= identk
    # data-parallel PPO on real rollouts
    from drloco_b200.ppo import PPO, PPOConfig
    from drloco_b200.vec_env import vec_env
    env = vec_env("StraightMimicWalker", num_envs=256, seed=33 + 100 * rank, device=f"cuda:{local}",
                  env_id_offset=rank * 256)
    cfg = PPOConfig(batch_size=256 * 16, minibatch_size=1024, total_steps=256 * 16 * world * 3)
    agent = PPO(env, cfg, seed=0)
    agent.learn(log_every=1)
    chk = agent.parameter_checksum()
    same = all(torch.equal(x, chk) for x in gathered(chk))
    out["ppo_replicas_bit_identical"] = same
    out["ppo_steps"] = agent.num_timesteps
    ok &= same and agent.num_timesteps == cfg.total_steps
    env.close()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
