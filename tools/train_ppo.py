#!/usr/bin/env python
"""Train the straight walker with PPO on GPU rollouts (mirror of reference drloco/train.py:77-139, next-tier demo).

Usage: python tools/train_ppo.py --envs 4096 --steps 16000000 --out gpurun_out/ppo_curve.json
       python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
              tools/train_ppo.py --envs 512 --steps 16000000 --out gpurun_out/ppo_8gpu.json     (data parallel)
       python tools/train_ppo.py --cpu-oracle --envs 8 --steps 1000000 --reference-hypers --out ...   (the same learner
              on the CPU oracle stack: the comparison curve of BASELINE.json configs[4])
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
    ap.add_argument("--cpu-oracle", action="store_true",
                    help="drive the learner with the CPU oracle stack (oracle/env_oracle.py) instead of the GPU env")
    ap.add_argument("--reference-hypers", action="store_true",
                    help="the reference's rollout shape: 16384 samples per update, minibatch 2048 (hypers.py:78-79)")
    ap.add_argument("--no-eval", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    if world > 1:
        torch.cuda.set_device(local)
        import datetime
        # a short collective timeout: a rank that falls out of step must fail the run quickly, not hold the GPUs
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local),
                                             timeout=datetime.timedelta(seconds=90))
    if args.cpu_oracle:
        from tools.oracle_tensor_env import OracleTensorEnv
        env = OracleTensorEnv("StraightMimicWalker", args.envs, seed=33 + args.seed)
    else:
        env = vec_env("StraightMimicWalker", num_envs=args.envs, seed=33 + args.seed + 100 * rank, norm_rew=True,
                      device=f"cuda:{local}", env_id_offset=rank * args.envs)
    cfg = PPOConfig(total_steps=args.steps)
    if args.reference_hypers:
        cfg.batch_size, cfg.minibatch_size = 4096 * 4, 512 * 4
    else:
        # the reference collects 16384 samples per update with 8 envs (2048 steps each); with thousands of envs keep
        # rollouts long enough for GAE to see consequences: 32 control steps per env per update
        cfg.batch_size = args.batch or args.envs * 32
        cfg.minibatch_size = args.minibatch or max(2048, cfg.batch_size // 8)
    agent = PPO(env, cfg, seed=args.seed)

    def cb(a, row):
        if rank == 0:
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
    if not args.cpu_oracle:
        torch.cuda.synchronize()
    wall = time.time() - t0
    chk = agent.parameter_checksum()
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        torch.distributed.all_gather(allc, chk)
        lockstep = all(torch.equal(c, allc[0]) for c in allc)
    else:
        lockstep = True
    ev = ev2 = None
    if rank == 0 and not args.no_eval and not args.cpu_oracle:
        ev = evaluate_walking(agent.policy, env)
        print("evaluation (deterministic policy, 20 deterministic inits):", json.dumps(ev), flush=True)
        ev2 = evaluate_walking(agent.policy, env, steady_state_counters=True)
        print("evaluation with the training-time desired-velocity counter (Q3):", json.dumps(ev2), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump({"envs_per_rank": args.envs, "world": world, "backend": "cpu-oracle" if args.cpu_oracle else "b200",
                       "total_steps": agent.num_timesteps, "wall_s": wall,
                       "env_steps_per_s_incl_learning": agent.num_timesteps / wall,
                       "replicas_bit_identical": lockstep,
                       "config": {k: v for k, v in vars(cfg).items()}, "curve": agent.log, "evaluation": ev,
                       "evaluation_steady_state_counters": ev2}, f, indent=1)
        print("done: %.1f s, %.2e env-steps/s incl. learning, replicas bit-identical: %s"
              % (wall, agent.num_timesteps / wall, lockstep))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
