"""PPO consumer of the GPU rollouts — SURVEY.md §8f "next" row 1 (beyond the hot-path scope; kept deliberately small).

Restates what reference ``drloco/train.py:77-139`` asks Stable-Baselines3 1.0 to do, with the reference's
hyper-parameters (``drloco/config/hypers.py:68-116``) and network (``drloco/custom/policies.py:13-80``: 2 x 512 tanh hidden
layers *shared* between policy and value head — Q15 —, state-independent log-std initialised at -0.75), on device
tensors end to end: observations, actions, rewards and the rollout buffer never leave HBM; the environment is stepped
through ``B200VecNormalize.step_tensor``.  The MLP uses plain PyTorch (cuBLAS GEMMs): it is not part of the hot path.

Data parallel over the GPUs of a node (one process per GPU, ``torch.distributed``): every rank rolls out its own shard
of the environments (no exchange inside ``step`` beyond the VecNormalize statistics), the learner replicas start from
rank 0's parameters and all-reduce ONE flat gradient bucket per minibatch (the parameter gradients are views of it), so
that the replicas stay bit-identical; episode statistics are all-reduced at the logging cadence.
"""
from __future__ import annotations

import dataclasses
import math
import time
from typing import Callable, List, Optional

import torch
import torch.nn as nn


@dataclasses.dataclass
class PPOConfig:
    """defaults = reference hypers.py (StraightMimicWalker, 200 Hz)."""
    gamma: float = 0.995                 # hypers.py:68
    gae_lambda: float = 0.95             # SB3 default
    batch_size: int = 4096 * 4           # hypers.py:79 samples per update (n_steps * n_envs)
    minibatch_size: int = 512 * 4        # hypers.py:78
    n_epochs: int = 4                    # hypers.py:116
    lr_start: float = 500e-6             # hypers.py:85-87, linear decay to lr_final
    lr_final: float = 1e-6
    clip_range: float = 0.15             # hypers.py:109 (policy and value function, train.py:114-115)
    ent_coef: float = -0.0075            # hypers.py:113
    vf_coef: float = 0.5                 # SB3 default
    max_grad_norm: float = 0.5           # SB3 default
    init_logstd: float = -0.75           # hypers.py:75
    hidden: tuple = (512, 512)           # hypers.py:97
    total_steps: int = int(8e6)          # hypers.py:90


class ActorCritic(nn.Module):
    """CustomActorCriticPolicy (policies.py:54-80): shared tanh trunk, linear action mean, linear value."""

    def __init__(self, obs_dim: int, act_dim: int, hidden=(512, 512), init_logstd=-0.75):
        super().__init__()
        layers, d = [], obs_dim
        for h in hidden:
            layers += [nn.Linear(d, h), nn.Tanh()]
            d = h
        self.trunk = nn.Sequential(*layers)
        self.action_net = nn.Linear(d, act_dim)
        self.value_net = nn.Linear(d, 1)
        self.log_std = nn.Parameter(torch.full((act_dim,), float(init_logstd)))
        for m in self.trunk:                              # SB3 ortho_init: sqrt(2) trunk, 0.01 policy, 1 value
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, math.sqrt(2))
                nn.init.zeros_(m.bias)
        nn.init.orthogonal_(self.action_net.weight, 0.01)
        nn.init.zeros_(self.action_net.bias)
        nn.init.orthogonal_(self.value_net.weight, 1.0)
        nn.init.zeros_(self.value_net.bias)

    def forward(self, obs):
        z = self.trunk(obs)
        return self.action_net(z), self.value_net(z).squeeze(-1)

    def dist(self, mean):
        return torch.distributions.Normal(mean, self.log_std.exp().expand_as(mean))


class PPO:
    """On-policy loop: collect n_steps x N transitions on the device, GAE(lambda), clipped surrogate updates."""

    def __init__(self, env, cfg: Optional[PPOConfig] = None, seed: int = 0, distributed: Optional[bool] = None):
        self.env, self.cfg = env, cfg or PPOConfig()
        self.device = env.device
        dist = torch.distributed
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.distributed = bool(distributed)
        self.world = dist.get_world_size() if self.distributed else 1
        self.rank = dist.get_rank() if self.distributed else 0
        torch.manual_seed(seed)
        venv = env.venv if hasattr(env, "venv") else env
        self.N, self.D, self.A = venv.num_envs, venv.obs_dim, venv.act_dim
        self.n_steps = max(1, self.cfg.batch_size // self.N)          # train.py:112
        self.policy = ActorCritic(self.D, self.A, self.cfg.hidden, self.cfg.init_logstd).to(self.device)
        # one flat gradient bucket: every parameter's .grad is a view of it (a single all-reduce per minibatch)
        params = list(self.policy.parameters())
        self._gradbuf = torch.zeros(sum(p.numel() for p in params), device=self.device)
        off = 0
        for p in params:
            p.grad = self._gradbuf[off:off + p.numel()].view_as(p)
            off += p.numel()
        if self.distributed:                                  # replicas start from rank 0's parameters
            for p in params:
                dist.broadcast(p.data, src=0)
        torch.manual_seed(seed + 7919 * self.rank)            # exploration noise differs between the ranks
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=self.cfg.lr_start, eps=1e-5)
        # minibatch permutations: the same seed on every rank (each rank permutes its own shard identically)
        self._perm_gen = torch.Generator(device=self.device)
        self._perm_gen.manual_seed(1_000_003 + seed)
        self.num_timesteps = 0
        self.step_callback: Optional[Callable[[], object]] = None    # e.g. TrainingMonitor.on_step
        self.log: List[dict] = []
        dev, T, N = self.device, self.n_steps, self.N
        self.buf = dict(obs=torch.zeros(T, N, self.D, device=dev), act=torch.zeros(T, N, self.A, device=dev),
                        logp=torch.zeros(T, N, device=dev), val=torch.zeros(T, N, device=dev),
                        rew=torch.zeros(T, N, device=dev), done=torch.zeros(T, N, device=dev))

    def save(self, path: str) -> None:
        """checkpoint of the learner in the layout of SB3's model zip (reference utils.save_model, utils.py:175-184:
        ``models/model_<ckpt>.zip``): policy state dict under SB3's key names, optimiser state, step counter and
        hyper-parameters in the json ``data`` entry.  The env statistics are saved by the env itself."""
        from .checkpoint import save_sb3_zip
        save_sb3_zip(path, self.policy, self.opt, data={"num_timesteps": int(self.num_timesteps),
                                                        "ppo_config": dataclasses.asdict(self.cfg)})

    def load(self, path: str) -> "PPO":
        import io
        import zipfile
        from .checkpoint import load_sb3_zip
        data = load_sb3_zip(path, self.policy, map_location=self.device)
        with zipfile.ZipFile(path) as z:
            if "policy.optimizer.pth" in z.namelist():
                osd = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), map_location=self.device,
                                 weights_only=True)
                if osd.get("param_groups"):
                    self.opt.load_state_dict(osd)
        self.num_timesteps = int(data.get("num_timesteps", self.num_timesteps))
        return self

    def _lr(self) -> float:                                            # schedules.py:16-33 LinearDecay
        frac = min(1.0, self.num_timesteps / max(1, self.cfg.total_steps))
        return self.cfg.lr_start + frac * (self.cfg.lr_final - self.cfg.lr_start)

    @torch.no_grad()
    def collect(self, obs):
        b = self.buf
        for t in range(self.n_steps):
            mean, val = self.policy(obs)
            d = self.policy.dist(mean)
            act = d.sample()
            b["obs"][t], b["act"][t], b["val"][t] = obs, act, val
            b["logp"][t] = d.log_prob(act).sum(-1)
            # the env clips to [-1, 1] itself (mimic_env.py:181); SB3's clip to the +-300 action space is a no-op
            nobs, rew, done = self.env.step_tensor(act.contiguous())
            b["rew"][t], b["done"][t] = rew, done.float()
            obs = nobs.clone()
            if self.step_callback is not None:                       # SB3 BaseCallback._on_step cadence
                self.step_callback()
        _, last_val = self.policy(obs)
        # GAE; SB3 1.0 does not bootstrap from the terminal observation
        adv = torch.zeros_like(b["rew"])
        last = torch.zeros(self.N, device=self.device)
        for t in reversed(range(self.n_steps)):
            nv = last_val if t == self.n_steps - 1 else b["val"][t + 1]
            nonterminal = 1.0 - b["done"][t]
            delta = b["rew"][t] + self.cfg.gamma * nv * nonterminal - b["val"][t]
            last = delta + self.cfg.gamma * self.cfg.gae_lambda * nonterminal * last
            adv[t] = last
        self.num_timesteps += self.n_steps * self.N * self.world      # global count: every rank collects its shard
        return obs, adv, adv + b["val"]

    def update(self, adv, ret):
        cfg, b = self.cfg, self.buf
        for g in self.opt.param_groups:
            g["lr"] = self._lr()
        flat = lambda x: x.reshape(-1, *x.shape[2:])                  # noqa: E731
        obs, act, logp0, val0, adv, ret = map(flat, (b["obs"], b["act"], b["logp"], b["val"], adv, ret))
        n = obs.shape[0]
        stats = []
        for _ in range(cfg.n_epochs):
            perm = torch.randperm(n, device=self.device, generator=self._perm_gen)
            for s in range(0, n, cfg.minibatch_size):
                idx = perm[s:s + cfg.minibatch_size]
                mean, val = self.policy(obs[idx])
                d = self.policy.dist(mean)
                logp = d.log_prob(act[idx]).sum(-1)
                a = adv[idx]
                a = (a - a.mean()) / (a.std() + 1e-8)
                ratio = (logp - logp0[idx]).exp()
                pl = -torch.min(a * ratio, a * ratio.clamp(1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
                vclip = val0[idx] + (val - val0[idx]).clamp(-cfg.clip_range, cfg.clip_range)     # clip_range_vf
                vl = ((ret[idx] - vclip) ** 2).mean()
                ent = d.entropy().sum(-1).mean()
                loss = pl + cfg.vf_coef * vl - cfg.ent_coef * ent
                self._gradbuf.zero_()                                 # grads are views of the bucket: keep them
                loss.backward()
                if self.distributed:                                  # mean gradient over the ranks' shards
                    torch.distributed.all_reduce(self._gradbuf, op=torch.distributed.ReduceOp.SUM)
                    self._gradbuf /= self.world
                nn.utils.clip_grad_norm_(self.policy.parameters(), cfg.max_grad_norm)
                self.opt.step()
                stats.append((pl.detach(), vl.detach(), ent.detach()))
        return [torch.stack(x).mean().item() for x in zip(*stats)]

    def learn(self, total_steps: Optional[int] = None, log_every: int = 10,
              callback: Optional[Callable[["PPO", dict], None]] = None):
        total = total_steps or self.cfg.total_steps
        self.cfg.total_steps = total
        venv = self.env.venv if hasattr(self.env, "venv") else self.env
        obs = self.env.reset_tensor().clone()
        it, t0 = 0, time.time()
        venv.reset_stats()
        while self.num_timesteps < total:
            obs, adv, ret = self.collect(obs)
            pl, vl, ent = self.update(adv, ret)
            it += 1
            if it % log_every == 0 or self.num_timesteps >= total:
                st = self._global_stats(venv)
                row = dict(steps=self.num_timesteps, wall_s=time.time() - t0,
                           mean_step_reward=self._global_mean_reward(),
                           mean_ep_len=st["ep_len_sum"] / max(1.0, st["episodes"]),
                           mean_ep_ret=st["ep_ret_sum"] / max(1.0, st["episodes"]),
                           moved_distance=st["moved_distance_sum"] / max(1.0, st["episodes"]),
                           pos_rew=st["pos_rew_sum"] / max(1.0, st["rew_steps"]),
                           vel_rew=st["vel_rew_sum"] / max(1.0, st["rew_steps"]),
                           episodes=st["episodes"], policy_loss=pl, value_loss=vl, entropy=ent, lr=self._lr())
                self.log.append(row)
                if callback:
                    callback(self, row)
        return self


    def _global_stats(self, venv) -> dict:
        """episode statistics since the last call, summed over the ranks (logging cadence)"""
        st = venv.stats()
        if self.distributed:
            keys = sorted(st)
            t = torch.tensor([float(st[k]) for k in keys], dtype=torch.float64, device=self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
            st = dict(zip(keys, t.cpu().tolist()))
        venv.reset_stats()
        return st

    def _global_mean_reward(self) -> float:
        if not hasattr(self.env, "get_original_reward"):
            return float("nan")
        m = self.env.get_original_reward().mean().reshape(1).clone()
        if self.distributed:
            torch.distributed.all_reduce(m, op=torch.distributed.ReduceOp.SUM)
            m /= self.world
        return m.item()

    def parameter_checksum(self) -> torch.Tensor:
        """float64 (sum, sum of squares) over all parameters: equal on every rank iff the replicas are in lockstep"""
        v = torch.cat([p.detach().reshape(-1) for p in self.policy.parameters()]).double()
        return torch.stack([v.sum(), (v * v).sum()])


@torch.no_grad()
def evaluate_walking(policy: ActorCritic, train_env, n_episodes: int = 20, min_stable_distance: float = 15.0,
                     steady_state_counters: bool = False):
    """Deterministic evaluation in the spirit of reference ``TrainingMonitor.eval_walking`` (callback.py:272-390):
    ``n_episodes`` episodes from the reference's deterministic initial states (evaluation mode,
    straight_walk_trajecs.py:237-265), deterministic actions, frozen normalisation statistics; reports walked distance,
    episode duration and the count of stable walks (callback.py:336-349).  The episodes run side by side in one small
    batched env instead of one after the other."""
    from .vec_env import B200MimicVecEnv, B200VecNormalize
    import numpy as np
    venv = train_env.venv
    spec = venv.spec
    t = spec.mocap
    env = B200MimicVecEnv(spec.cfg.env_id, num_envs=n_episodes, device=venv.device, cfg=spec.cfg, spec=spec, seed=0)
    # a local, single-process wrapper: the evaluation may run on one rank only (no collective at construction)
    vn = B200VecNormalize(env, training=False, norm_reward=False, distributed=False)
    vn.load_state_dict({**train_env.state_dict(), "norm_reward": False})
    # the reference plays EVAL_N_TIMES consecutive episodes in one env in evaluation mode (callback.py:286-297); here
    # env i plays episode i: deterministic init i, including the reference's table aliasing
    counters = np.arange(n_episodes) % spec.cfg.eval_n_times
    env.env_method("activate_evaluation")
    if steady_state_counters:
        # count_steps_same_vel is never reset in the reference (Q3): after ~30 mocap steps of an env's lifetime the
        # desired-velocity observation is stuck at step_vel[0], which is what the policy saw for nearly all of its
        # training.  A freshly built evaluation env starts the counter at 1 and feeds step_vel[i_step] instead; this
        # switch reproduces the training-time (steady-state) counter in the evaluation env.
        q, v, cur = env.get_state()
        cur[:, 2] = 10 ** 6
        env.set_state(None, None, cur)
    env.set_det_init_counters(counters)
    obs = vn.reset_tensor().clone()
    n = n_episodes
    alive = torch.ones(n, dtype=torch.bool, device=venv.device)
    ep_len = torch.zeros(n, device=venv.device)
    rew_sum = torch.zeros(n, device=venv.device)
    dist = torch.zeros(n, device=venv.device)
    for _ in range(spec.cfg.ep_dur_max + 1):
        mean, _ = policy(obs)
        nobs, rew, done = vn.step_tensor(mean.contiguous())
        d = done.bool()
        ex = env.extras()
        # walked distance (mimic_env.py:295).  The reference reads it only on steps that did not end the episode
        # ("when done=True is returned, the env is already resetted", callback.py:313-316), so an episode's distance
        # is the value one step before its end: `dist` simply stops being updated on the done step.
        ep_len += alive.float()
        rew_sum += torch.where(alive & ~d, rew, torch.zeros_like(rew))
        dist = torch.where(alive & ~d, ex[:, 3], dist)
        alive &= ~d
        obs = nobs.clone()
        if not bool(alive.any()):
            break
    env.close()
    ep_len_h, dist_h = ep_len.cpu().numpy(), dist.cpu().numpy()
    mean_rew = (rew_sum / torch.clamp(ep_len - 1, min=1)).cpu().numpy()
    reached = int((dist_h >= min_stable_distance).sum())
    no_fall = int(((ep_len_h >= spec.cfg.ep_dur_max) & (dist_h >= 0.5 * min_stable_distance)).sum())
    return dict(mean_walked_distance=float(dist_h.mean()), min_walked_distance=float(dist_h.min()),
                mean_episode_duration=float(ep_len_h.mean() / spec.cfg.ep_dur_max),
                min_episode_duration=float(ep_len_h.min()),
                mean_walking_speed=float((dist_h / (ep_len_h / spec.cfg.ctrl_freq)).mean()),
                mean_reward_means=float((mean_rew.mean() - spec.cfg.alive_bonus) / spec.cfg.rew_scale),
                count_stable_walks=max(reached, no_fall), n_episodes=n_episodes,
                # per-episode values, as TrainingMonitor.eval_walking collects them (callback.py:277,306-311)
                moved_distances=dist_h.tolist(), ep_durs=ep_len_h.tolist(), mean_rewards=mean_rew.tolist())
