"""Data-parallel PPO learner on CPU (gloo, world_size 2): the replicas see different rollouts, all-reduce one flat gradient
bucket per minibatch and must stay bit-identical (drloco_b200/ppo.py; reference train.py:110-133 is single-process)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeVenv:
    """stand-in for B200VecNormalize on CPU tensors: linear dynamics, reward = -|obs|, random episode ends"""

    def __init__(self, n, d, a, seed):
        self.num_envs, self.obs_dim, self.act_dim = n, d, a
        self.device = torch.device("cpu")
        self.g = torch.Generator().manual_seed(seed)
        self.obs = torch.zeros(n, d)
        self.W = torch.randn(a, d, generator=torch.Generator().manual_seed(0)) * 0.1
        self._ep = 0.0

    def reset_tensor(self):
        self.obs = torch.randn(self.num_envs, self.obs_dim, generator=self.g)
        return self.obs

    def step_tensor(self, act):
        self.obs = 0.9 * self.obs + act.clamp(-1, 1) @ self.W + 0.05 * torch.randn(self.obs.shape, generator=self.g)
        rew = -self.obs.abs().mean(1)
        done = (torch.rand(self.num_envs, generator=self.g) < 0.02).to(torch.uint8)
        self._ep += float(done.sum())
        return self.obs, rew, done

    def stats(self):
        return dict(episodes=self._ep, ep_len_sum=50.0 * self._ep, ep_ret_sum=0.0, moved_distance_sum=0.0,
                    pos_rew_sum=0.0, vel_rew_sum=0.0, rew_steps=1.0)

    def reset_stats(self):
        self._ep = 0.0


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from drloco_b200.ppo import PPO, PPOConfig
    env = _FakeVenv(16, 6, 3, seed=100 + rank)                   # every rank rolls out its own shard
    cfg = PPOConfig(batch_size=16 * 8, minibatch_size=32, n_epochs=2, hidden=(32, 32), total_steps=16 * 8 * 2 * 3)
    agent = PPO(env, cfg, seed=3)
    assert agent.distributed and agent.world == world
    c0 = agent.parameter_checksum()
    agent.learn(log_every=1)
    c1 = agent.parameter_checksum()
    obs_sum = float(agent.buf["obs"].double().sum())
    out[rank] = (c0.tolist(), c1.tolist(), obs_sum, agent.num_timesteps, [r["episodes"] for r in agent.log])
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_replicas_stay_identical():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (a0, a1, sa, na, ea), (b0, b1, sb, nb, eb) = out[0], out[1]
    assert a0 == b0                                              # same start (rank 0's parameters were broadcast)
    assert a1 == b1 and a1 != a0                                 # bit-identical after three updates, and they did move
    assert sa != sb                                              # although the rollouts differed between the ranks
    assert na == nb == 16 * 8 * 2 * 3                            # global step count
    assert ea == eb and sum(ea) > 0                              # episode statistics all-reduced (same on both ranks)
