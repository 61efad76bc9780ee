#!/usr/bin/env python
"""Train the straight walker with PPO on GPU rollouts (mirror of reference drloco/train.py:77-139, next-tier demo).

Usage: python tools/train_ppo.py --envs 4096 --steps 16000000 --out gpurun_out/ppo_curve.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from drloco_b200.ppo import PPO, PPOConfig, evaluate_walking  # noqa: E402
from drloco_b200.vec_env import vec_env  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=int(16e6))
    ap.add_argument("--batch", type=int, default=0, help="samples per update (default: 32 steps per env)")
    ap.add_argument("--minibatch", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/ppo_curve.json")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--monitor", default="", help="directory: attach the TrainingMonitor (evaluation, checkpoints, "
                                                  "scalars with the reference's tag names) and write there")
    args = ap.parse_args()
    env = vec_env("StraightMimicWalker", num_envs=args.envs, seed=33 + args.seed, norm_rew=True)
    cfg = PPOConfig(total_steps=args.steps)
    # the reference collects 16384 samples per update with 8 envs (2048 steps each); with thousands of envs keep
    # rollouts long enough for GAE to see consequences: 32 control steps per env per update
    cfg.batch_size = args.batch or args.envs * 32
    cfg.minibatch_size = args.minibatch or max(2048, cfg.batch_size // 8)
    agent = PPO(env, cfg, seed=args.seed)

    def cb(a, row):
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)

    mon = None
    if args.monitor:
        from drloco_b200.training_monitor import TrainingMonitor
        mon = TrainingMonitor(agent, env.venv.cfg, args.monitor.rstrip("/") + "/", verbose=1)
        mon.on_training_start()
        agent.step_callback = mon.on_step
    t0 = time.time()
    agent.learn(args.steps, log_every=5, callback=cb)
    if mon:
        mon.on_training_end()
        print("checkpoints kept:", mon.saved, " steps to convergence:", mon.steps_to_convergence, flush=True)
    torch.cuda.synchronize()
    ev = evaluate_walking(agent.policy, env)
    print("evaluation (deterministic policy, 20 deterministic inits):", json.dumps(ev), flush=True)
    ev2 = evaluate_walking(agent.policy, env, steady_state_counters=True)
    print("evaluation with the training-time desired-velocity counter (Q3):", json.dumps(ev2), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"envs": args.envs, "total_steps": agent.num_timesteps, "wall_s": time.time() - t0,
                   "config": {k: v for k, v in vars(cfg).items()}, "curve": agent.log, "evaluation": ev, "evaluation_steady_state_counters": ev2}, f, indent=1)
    print("done: %.1f s, %.2e env-steps/s incl. learning" % (time.time() - t0, agent.num_timesteps / (time.time() - t0)))


if __name__ == "__main__":
    main()
