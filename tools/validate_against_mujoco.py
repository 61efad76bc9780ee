#!/usr/bin/env python
"""Harness to pin the physics oracle (and through it the CUDA path) against a real MuJoCo — for machines that have one.

The build image has neither ``mujoco`` nor ``mujoco_py``, so oracle/walker_physics.c is a restatement
of MuJoCo's documented pipeline that could not be compared with MuJoCo itself: PARITY UNPINNED.  This script is the
comparison, ready to run wherever ``pip install mujoco`` works and a DRLoco checkout is at hand.  It has NOT been
executed in the build image.

    python tools/validate_against_mujoco.py --reference /path/to/DRLoco [--env StraightMimicWalker] [--steps 20]

What it compares, from identical states and controls (float64 both sides):
  1. model constants MuJoCo compiles: body_invweight0, dof_invweight0, total mass;
  2. one forward evaluation at random states: mass matrix (mj_fullM), bias force (qfrc_bias), number of contacts,
     constrained acceleration qacc (Newton solver, tolerance tightened to 1e-12);
  3. rollouts of frame_skip x RK4 steps with random torques: per-control-step max |dqpos|, |dqvel|.
Exit code 0 when every check is within the printed tolerance.
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

XML = {"StraightMimicWalker": "drloco/mujoco/xml/walker3d_flat_feet.xml",
       "MimicWalker165cm65kg": "drloco/mujoco/xml/walker_165cm_65kg.xml"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="path of a rgalljamov/DRLoco checkout")
    ap.add_argument("--env", default="StraightMimicWalker", choices=sorted(XML))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    try:
        import mujoco
    except ImportError:
        print("the `mujoco` package is not installed here: nothing to compare against (parity stays unpinned)")
        return 2
    from drloco_b200.config import EnvConfig
    from drloco_b200.model import get_model
    from drloco_b200.walkers import make_spec
    from oracle.physics import OraclePhysics

    m = get_model(args.env)
    spec = make_spec(EnvConfig(env_id=args.env))
    P = OraclePhysics(m)
    mjm = mujoco.MjModel.from_xml_path(os.path.join(args.reference, XML[args.env]))
    mjm.opt.tolerance = 1e-12
    mjm.opt.iterations = 200
    mjd = mujoco.MjData(mjm)
    rng = np.random.default_rng(args.seed)
    ok = True

    def check(name, got, want, tol):
        nonlocal ok
        err = float(np.max(np.abs(np.asarray(got) - np.asarray(want)) / np.maximum(1.0, np.abs(np.asarray(want)))))
        flag = "ok " if err <= tol else "FAIL"
        ok = ok and err <= tol
        print(f"[{flag}] {name:32s} max rel err {err:.3e} (tol {tol:.0e})")

    # 1. compile-time constants (MuJoCo body 0 is the world)
    check("total mass", m.total_mass, float(np.sum(mjm.body_mass)), 1e-12)
    check("dof_invweight0", m.dof_invweight0, mjm.dof_invweight0, 1e-9)
    check("body_invweight0", m.body_invweight0, mjm.body_invweight0[1:], 1e-9)
    check("qpos0", m.qpos0, mjm.qpos0, 1e-12)

    # 2. forward evaluations near the mocap poses, with and without ground contact
    t = spec.mocap
    nv, nu = m.nv, m.nu
    for k in range(8):
        row = rng.integers(0, t.n_samples)
        q = t.ref[row, :nv].copy()
        v = t.ref[row, nv:2 * nv].copy() + 0.3 * rng.standard_normal(nv)
        q[3:] += 0.05 * rng.standard_normal(nv - 3)
        q[2] -= P.site_xpos(q)[:, 2].min() + (0.003 if k % 2 else -0.05)      # odd k: 3 mm penetration
        ctrl = rng.uniform(-300, 300, nu)
        mjd.qpos[:], mjd.qvel[:], mjd.ctrl[:] = q, v, ctrl
        mjd.qacc_warmstart[:] = 0
        mujoco.mj_forward(mjm, mjd)
        Mfull = np.zeros((nv, nv))
        mujoco.mj_fullM(mjm, Mfull, mjd.qM)
        a, diag = P.forward(q, v, ctrl)
        check(f"state {k}: mass matrix", P.mass_matrix(q), Mfull, 1e-10)
        check(f"state {k}: bias force", P.bias(q, v), mjd.qfrc_bias, 1e-9)
        print(f"         contacts: oracle {diag.ncon}  mujoco {mjd.ncon};  constraint rows: {diag.nefc} / {mjd.nefc}")
        ok = ok and diag.ncon == mjd.ncon
        check(f"state {k}: qacc", a, mjd.qacc, 1e-6)

    # 3. rollouts (control held for frame_skip RK4 steps, like MujocoEnv.do_simulation)
    fs = spec.frame_skip
    row = rng.integers(0, t.n_samples)
    q = t.ref[row, :nv].copy()
    v = t.ref[row, nv:2 * nv].copy()
    q[2] -= P.site_xpos(q)[:, 2].min()
    mujoco.mj_resetData(mjm, mjd)
    mjd.qpos[:], mjd.qvel[:] = q, v
    qo, vo = q.copy(), v.copy()
    P.qacc_warm[:] = 0
    for k in range(args.steps):
        ctrl = rng.uniform(-300, 300, nu)
        mjd.ctrl[:] = ctrl
        for _ in range(fs):
            mujoco.mj_step(mjm, mjd)
        P.step(qo, vo, ctrl, fs)
        dq, dv = np.abs(qo - mjd.qpos).max(), np.abs(vo - mjd.qvel).max()
        print(f"control step {k:3d}: max |dqpos| {dq:.3e}  max |dqvel| {dv:.3e}")
        if k < 5:
            ok = ok and dq < 1e-6 and dv < 1e-4
    print("RESULT:", "oracle matches MuJoCo within the stated tolerances" if ok else "MISMATCH - see FAIL lines")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
