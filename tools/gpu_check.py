#!/usr/bin/env python
"""Developer diagnostic (run on a GPU box): prints GPU-vs-oracle discrepancies stage by stage instead of asserting."""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from drloco_b200.config import EnvConfig  # noqa: E402
from drloco_b200.vec_env import B200MimicVecEnv  # noqa: E402
from drloco_b200.walkers import make_spec  # noqa: E402
from oracle.env_oracle import OracleVecEnv  # noqa: E402
from oracle.physics import OraclePhysics  # noqa: E402

np.set_printoptions(precision=5, suppress=True, linewidth=220)


def forward_pieces(env_id="StraightMimicWalker", n=32, seed=0):
    """one dynamics evaluation (Euler, frame_skip=1, dump on): mass matrix, bias, qacc vs the float64 oracle."""
    cfg = EnvConfig(env_id=env_id, integrator="euler")
    spec = make_spec(cfg)
    env = B200MimicVecEnv(env_id, num_envs=n, cfg=cfg, spec=spec)
    env.debug_set(frame_skip_override=1, enable_dump=True)
    rng = np.random.default_rng(seed)
    m = spec.model
    nv, nu = m.nv, m.nu
    t = spec.mocap
    rows = rng.integers(0, t.n_samples, n)
    q = t.ref[rows, :nv].copy()
    v = t.ref[rows, nv:2 * nv].copy()
    # put the walker near the ground (some in the air, some penetrating), perturb joints
    P = OraclePhysics(m)
    for i in range(n):
        q[i, 3:] += 0.1 * rng.standard_normal(nv - 3)
        q[i, 0] = rng.uniform(-1, 30)
        low = P.site_xpos(q[i])[:, 2].min()
        q[i, 2] -= low + rng.uniform(-0.004, 0.004)
        v[i] += 0.3 * rng.standard_normal(nv)
    env.reset()
    cur = np.zeros((n, 4), np.int32)
    cur[:, 2] = 1
    env.set_state(q, v, cur)
    a = rng.uniform(-1, 1, (n, nu)).astype(np.float32)
    env.step(a)
    d = env.debug_read()
    worst = dict(M=0, c=0, a=0)
    for i in range(n):
        ctrl = (a[i] * 300).astype(np.float64)
        left = spec.mirror and bool(t.left_step[0])
        if left:
            oi, osn, ai, asn = spec.mirror_tables()
            ctrl = ctrl[ai] * asn
        qq, vv = q[i].astype(np.float32).astype(np.float64), v[i].astype(np.float32).astype(np.float64)
        Mo = P.mass_matrix(qq)
        co = P.bias(qq, vv)
        ao, dg = P.forward(qq, vv, ctrl)
        Mg = d[i, 2:2 + nv, :nv]          # rows r, lanes c  -> M[r][c]
        cg = d[i, 0, :nv]
        ag = d[i, 2 + nv, :nv]
        eM = np.abs(Mg - Mo).max() / np.abs(Mo).max()
        ec = np.abs(cg - co).max() / max(1.0, np.abs(co).max())
        ea = np.abs(ag - ao).max() / max(1.0, np.abs(ao).max())
        worst["M"], worst["c"], worst["a"] = max(worst["M"], eM), max(worst["c"], ec), max(worst["a"], ea)
        if i < 4 or ea > 1e-3:
            print(f"env {i}: ncon oracle {dg.ncon} gpu {int(d[i, 4 + nv, :].sum())}  relerr M {eM:.2e} c {ec:.2e} qacc {ea:.2e}"
                  f"  |a|max {np.abs(ao).max():.1f} zO {d[i, 3 + nv, 0]:.4f}")
            if ea > 1e-3:
                print("   a gpu", ag)
                print("   a ora", ao)
    print("forward pieces worst rel err:", worst)
    env.close()


def rollout(env_id="StraightMimicWalker", n=64, steps=40, seed=1, integrator="rk4"):
    cfg = EnvConfig(env_id=env_id, integrator=integrator)
    spec = make_spec(cfg)
    env = B200MimicVecEnv(env_id, num_envs=n, cfg=cfg, spec=spec)
    from drloco_b200 import cabi
    ora = OracleVecEnv(spec, n, lambda: OraclePhysics(spec.model, cabi.INTEGRATOR_RK4 if integrator == "rk4"
                                                      else cabi.INTEGRATOR_EULER))
    rng = np.random.default_rng(seed)
    t = spec.mocap
    istep = rng.integers(0, t.n_steps, n).astype(np.int32)
    pos = np.array([rng.integers(0, t.step_len[i]) for i in istep], np.int32)
    og = env.reset(inject=(istep, pos))
    oo = ora.reset(istep, pos)
    print("reset obs max err", np.abs(og - oo).max())
    ndone = 0
    for k in range(steps):
        a = rng.uniform(-1, 1, (n, spec.act_dim)).astype(np.float32)
        og, rg, dg, _ = env.step(a, inject=(istep, pos))
        oo, ro, do, _ = ora.step(a, istep, pos)
        qg, vg, cg = env.get_state()
        qo = np.stack([e.env.qpos for e in ora.envs])
        vo = np.stack([e.env.qvel for e in ora.envs])
        co = np.array([[e.env.refs.i_step, e.env.refs.pos, e.env.refs.count_steps_same_vel, e.env.ep_dur] for e in ora.envs])
        ndone += int(do.sum())
        print(f"step {k:3d}: done match {np.array_equal(dg, do)} ({int(do.sum())})  cursor match {np.array_equal(cg, co)}"
              f"  q err {np.abs(qg - qo).max():.2e}  v err {np.abs(vg - vo).max():.2e}  obs err {np.abs(og - oo).max():.2e}"
              f"  rew err {np.abs(rg - ro).max():.2e}")
        if not np.array_equal(dg, do):
            bad = np.nonzero(dg != do)[0]
            print("   done mismatch envs", bad, "qz gpu", qg[bad, 2], "oracle", qo[bad, 2])
            break
    print("episodes finished:", ndone, " stats:", env.stats())
    env.close()


def speed(n=4096, steps=50, env_id="StraightMimicWalker", integrator="rk4", lanes=0, block=0):
    cfg = EnvConfig(env_id=env_id, integrator=integrator)
    env = B200MimicVecEnv(env_id, num_envs=n, cfg=cfg, lanes_per_env=lanes)
    if block:
        env.debug_set(block_threads=block)
    env.reset_tensor()
    acts = torch.rand(16, n, env.act_dim, device="cuda") * 2 - 1
    for k in range(10):
        env.step_tensor(acts[k % 16])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        env.step_tensor(acts[k % 16])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = env.stats()
    print(f"{env_id} {integrator} n={n} lanes={env.launch_info()}: {ms:.3f} ms/step -> {n / ms * 1e3:.3e} env-steps/s;"
          f" episodes {st['episodes']:.0f} iters/eval {st['solver_iters'] / max(1, st['dyn_evals']):.2f}")
    env.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["pieces", "rollout", "speed"]
    t0 = time.time()
    if "pieces" in what:
        forward_pieces()
    if "rollout" in what:
        rollout()
    if "euler" in what:
        rollout(integrator="euler", steps=20)
    if "speed" in what:
        for n in (4096, 16384, 65536):
            speed(n)
        speed(4096, lanes=32)
        speed(4096, integrator="euler")
    if "w165" in what:
        forward_pieces("MimicWalker165cm65kg")
        rollout("MimicWalker165cm65kg", steps=20)
        speed(4096, env_id="MimicWalker165cm65kg")
    print("total", time.time() - t0)
