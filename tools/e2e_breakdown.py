#!/usr/bin/env python
"""Where the end-to-end (numpy API) step time goes: host enqueue / device chain / host post-processing (developer tool)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drloco_b200.vec_env import B200MimicVecEnv, B200VecNormalize  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    env = B200MimicVecEnv("StraightMimicWalker", num_envs=n)
    out = {}
    for graph in (True, False):
        for copy in (True, False):
            vn = B200VecNormalize(env)
            vn.use_graph, vn.copy_outputs = graph, copy
            vn.reset()
            acts = [np.random.default_rng(k).uniform(-1, 1, (n, 8)).astype(np.float32) for k in range(8)]
            for k in range(10):
                vn.step(acts[k % 8])
            ta = tw = tp = 0.0
            K = 200
            t_all0 = time.perf_counter()
            for k in range(K):
                t0 = time.perf_counter()
                vn.step_async(acts[k % 8])
                t1 = time.perf_counter()
                torch.cuda.current_stream().synchronize()
                t2 = time.perf_counter()
                vn.step_wait()
                t3 = time.perf_counter()
                ta += t1 - t0; tw += t2 - t1; tp += t3 - t2
            tot = time.perf_counter() - t_all0
            out[f"graph={graph},copy={copy}"] = dict(us_per_step=1e6 * tot / K, enqueue_us=1e6 * ta / K,
                                                     wait_us=1e6 * tw / K, post_us=1e6 * tp / K,
                                                     env_steps_per_s=n * K / tot)
            env._lib.drl_attach_vecnorm(env._handle, None, 0.0, None)
    # device-only reference: tensor API, serialised
    vn = B200VecNormalize(env)
    a = torch.rand(n, 8, device="cuda") * 2 - 1
    vn.reset_tensor()
    for k in range(10):
        vn.step_tensor(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(200):
        vn.step_tensor(a)
    torch.cuda.synchronize()
    out["tensor_api_serialized_us"] = 1e6 * (time.perf_counter() - t0) / 200
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
