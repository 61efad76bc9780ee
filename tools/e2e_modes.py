"""numpy-API end-to-end rate per rank for host_outputs = mapped | copy, alternating A B A B inside one launch
(python tools/e2e_modes.py, or under torchrun for N ranks).  Prints one JSON line on rank 0: per mode and repetition the
env-steps/s of the slowest rank and the per-rank step times."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch.distributed as dist
    from datetime import timedelta
    from drloco_b200.vec_env import B200MimicVecEnv, B200VecNormalize
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", timeout=timedelta(seconds=90), device_id=torch.device(f"cuda:{local}"))
    n, steps = 4096, int(os.environ.get("E2E_STEPS", "300"))
    acts = [np.random.default_rng(rank * 100 + k).uniform(-1, 1, (n, 8)).astype(np.float32) for k in range(8)]
    res = []
    for rep in range(2):
        for mode in ("mapped", "copy"):
            env = B200MimicVecEnv("StraightMimicWalker", num_envs=n, device=f"cuda:{local}", seed=rank,
                                  env_id_offset=rank * n)
            vn = B200VecNormalize(env, distributed=world > 1)
            vn.host_outputs = mode
            vn.reset()
            for k in range(20):
                vn.step(acts[k % 8])
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for k in range(steps):
                o, r, d, infos = vn.step(acts[k % 8])
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64)
            ts = [torch.zeros_like(t) for _ in range(world)]
            if world > 1:
                dist.all_gather(ts, t)
            else:
                ts = [t]
            per_rank_us = [float(x) / steps * 1e6 for x in ts]
            res.append(dict(mode=mode, rep=rep, env_steps_per_s=world * n * steps / max(float(x) for x in ts),
                            step_us_per_rank=[round(u, 1) for u in per_rank_us]))
            vn.close()
    if rank == 0:
        print(json.dumps(dict(world=world, steps=steps, cores=os.cpu_count(), torch_threads=torch.get_num_threads(),
                              runs=res)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
