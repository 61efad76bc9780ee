#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the reference's OWN environment code (SURVEY.md §8c recipe).

Runs only in the build container (needs /root/reference).  The reference's MimicWalker3dEnv / MimicEnv / Monitor /
StraightWalkingTrajectories classes are imported *unmodified* from /root/reference; the third-party packages they
import but that are not installed (gym, mujoco_py, seaborn, matplotlib, wandb, stable_baselines3) are replaced by
import stubs, and gym's ``MujocoEnv`` by a minimal stand-in whose ``sim`` is the float64 physics oracle
(oracle/walker_physics.c).  Two textual substitutions are applied while loading reference modules, both forced by
the checkout rather than chosen: ``PATH_REF_TRAJECS = PATH_CONSTANT_SPEED`` (the default ramp mocap is missing, Q8)
and ``from collections import Iterable`` -> ``collections.abc`` (Python >= 3.10).

Conditions of the run (recorded in the fixture):
  * smoothing state (drloco.common.utils._exp_weighted_averages) is swapped per env, i.e. SubprocVecEnv semantics (Q17);
  * the mocap array is restored to its pristine copy before every reset, i.e. the in-place accumulation of
    adjust_COM_Z_pos across episodes (Q4) is waived, as documented in DESIGN.md;
  * envs are stepped with DummyVecEnv semantics written out by hand (SB3 is not installed): on done keep the terminal
    observation and reset.

Usage: python tools/gen_golden.py            (writes tests/golden/w3d_rollout.npz, w3d_cursor.npz, w3d_eval.npz)
       python tools/gen_golden.py w165       (writes tests/golden/w165_rollout.npz; own process: other ENV_ID)
"""
import collections
import collections.abc
import copy
import os
import random
import sys
import tempfile
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from drloco_b200.model import get_model            # noqa: E402
from oracle.physics import OraclePhysics           # noqa: E402


# the import recipe (stubs, the MujocoEnv stand-in over the physics oracle, load_reference) is shared with the CPU arm of
# bench.py: baseline/ref_runner.py
from baseline.ref_runner import (FakeMujocoEnv, MujocoException, REF, _load_module_with, install_stubs,   # noqa: E402,F401
                                 load_reference)


def load_reference_w165(project_dir):
    """The reference configured for MimicWalker165cm65kg.  Substitutions, all of them settings the reference expects its
    user to edit: ENV_ID (config.py:18), policy mirroring off (hypers.py:23,31-39: "only works with the Straight
    Walker"), and the project path the loco3d loader reads its (missing) .mat from."""
    sys.path.insert(0, REF)
    work = os.path.join(tempfile.mkdtemp(), "code", "torch")
    os.makedirs(work)
    os.chdir(work)
    import torch  # noqa: F401
    install_stubs()
    import drloco.config  # noqa: F401
    _load_module_with("drloco.config.config", "drloco/config/config.py",
                      [("ENV_ID = 'StraightMimicWalker'", "ENV_ID = 'MimicWalker165cm65kg'")])
    _load_module_with("drloco.config.hypers", "drloco/config/hypers.py",
                      [("modifications_list = [MOD_CUSTOM_POLICY, MOD_MIRR_POLICY]",
                        "modifications_list = [MOD_CUSTOM_POLICY]")])
    import drloco.ref_trajecs  # noqa: F401
    _load_module_with("drloco.ref_trajecs.base_ref_trajecs", "drloco/ref_trajecs/base_ref_trajecs.py",
                      [("from collections import Iterable", "from collections.abc import Iterable")])
    import drloco.ref_trajecs.loco3d_trajecs as l3
    l3.get_project_path = lambda: project_dir.rstrip("/") + "/"
    from drloco.mujoco.mimic_walker_165cm_65kg import MimicWalker165cm65kgEnv
    from drloco.mujoco.monitor_wrapper import Monitor
    from drloco.common import utils
    return MimicWalker165cm65kgEnv, Monitor, utils


def gen_w165_rollout(n_envs=6, n_steps=160, seed=0, out="w165_rollout.npz"):
    """reference MimicWalker165cm65kgEnv (wrap cursor, joint-phase estimates, 2-D desired velocity, no mirroring) over
    the oracle physics, reading the synthetic loco3d recording (drloco_b200.ref_trajecs.loco3d_trajecs.synthetic_loco3d,
    seed 0) from a .mat file with the reference's schema (loco3d_trajecs.py:35-46)."""
    import scipy.io as spio
    from drloco_b200.ref_trajecs.loco3d_trajecs import synthetic_loco3d, N_ROWS
    proj = tempfile.mkdtemp()
    os.makedirs(os.path.join(proj, "mocaps/loco3d"))
    ang, vel = synthetic_loco3d()
    spio.savemat(os.path.join(proj, "mocaps/loco3d/loco3d_guoping.mat"),
                 {"angJoi": ang, "angDJoi": vel, "rowNameIK": np.array([f"row{i}" for i in range(N_ROWS)], dtype=object)})
    Env, Monitor, utils = load_reference_w165(proj)
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    envs, ewa = [], []
    for i in range(n_envs):
        utils._exp_weighted_averages = {}
        envs.append(Monitor(Env()))
        ewa.append({})
    pristine = [e.env.refs._qpos_full.copy() for e in envs]
    nv, nu = 19, 13
    D = envs[0].env.observation_space.shape[0]
    assert D == 47
    rsi_log = []
    for i, mon in enumerate(envs):
        refs = mon.env.refs
        orig = refs.get_random_init_state

        def wrapped(orig=orig, refs=refs):
            out_ = orig()
            rsi_log.append(refs._pos)
            return out_
        refs.get_random_init_state = wrapped
    T = n_steps
    g = dict(actions=np.zeros((T, n_envs, nu), np.float32), obs=np.zeros((T, n_envs, D)), rew=np.zeros((T, n_envs)),
             done=np.zeros((T, n_envs), np.uint8), terminal_obs=np.full((T, n_envs, D), np.nan),
             qpos=np.zeros((T, n_envs, nv)), qvel=np.zeros((T, n_envs, nv)), cursor=np.zeros((T, n_envs, 2), np.int32),
             ctrl=np.zeros((T, n_envs, nu)), comps=np.zeros((T, n_envs, 3)), walked=np.zeros((T, n_envs)),
             des_vel=np.zeros((T, n_envs, 2)), rsi=np.full((T + 1, n_envs), -1, np.int32))
    obs0 = np.zeros((n_envs, D))
    # RSI draws: mostly random, two envs close to the end of the recording so that the cursor wraps (base:100-103)
    L = envs[0].env.refs._trajec_len
    forced = {0: L - 40, 1: L - 7}
    for i in range(n_envs):
        utils._exp_weighted_averages = ewa[i]
        envs[i].env.refs._qpos_full[...] = pristine[i]           # Q4 waiver: no accumulation across episodes
        if i in forced:
            real = np.random.randint
            np.random.randint = lambda lo, hi, _v=forced[i]: _v
            obs0[i] = envs[i].env.reset()
            np.random.randint = real
        else:
            obs0[i] = envs[i].env.reset()
        g["rsi"][0, i] = rsi_log[-1]
    g["obs0"] = obs0
    g["qpos0"] = np.stack([e.env.sim.data.qpos.copy() for e in envs])
    g["qvel0"] = np.stack([e.env.sim.data.qvel.copy() for e in envs])
    g["cursor0"] = np.array([[e.env.refs._pos, e.env.ep_dur] for e in envs], np.int32)
    for t in range(T):
        a = rng.uniform(-1.3, 1.3, size=(n_envs, nu)).astype(np.float32)
        a[: n_envs // 2] *= 0.1
        g["actions"][t] = a
        for i, mon in enumerate(envs):
            utils._exp_weighted_averages = ewa[i]
            e = mon.env
            o, r, d, _ = mon.step(a[i])
            g["qpos"][t, i], g["qvel"][t, i] = e.sim.data.qpos, e.sim.data.qvel
            g["cursor"][t, i] = (e.refs._pos, e.ep_dur)
            g["ctrl"][t, i] = e.sim.data.ctrl
            g["comps"][t, i] = (e.pos_rew, e.vel_rew, e.com_rew)
            g["walked"][t, i] = e.walked_distance
            g["des_vel"][t, i] = e.desired_walking_speed
            g["rew"][t, i], g["done"][t, i] = r, d
            if d:
                g["terminal_obs"][t, i] = o
                e.refs._qpos_full[...] = pristine[i]
                o = e.reset()
                g["rsi"][t + 1, i] = rsi_log[-1]
            g["obs"][t, i] = o
    for name in ("ep_len_smoothed", "ep_ret_smoothed", "mean_reward_smoothed", "moved_distance",
                 "mean_ep_pos_rew_smoothed", "mean_ep_vel_rew_smoothed", "mean_ep_com_rew_smoothed",
                 "mean_abs_ep_torque_smoothed", "median_abs_torque_smoothed"):
        g["mon_" + name] = np.array([float(getattr(m, name)) for m in envs])
    g["mon_ep_lens_flat"] = np.array([x for m in envs for x in m.ep_lens], np.int32)
    g["meta"] = np.array("reference MimicWalker165cm65kgEnv+Monitor (ENV_ID and mirroring set in the config as the "
                         "reference asks its user to) over oracle physics on synthetic_loco3d(seed 0); Q4 waived; "
                         "per-env smoothing dicts; seed=%d" % seed)
    np.savez_compressed(os.path.join(REPO, "tests/golden", out), **g)
    print(out, "episodes:", int(g["done"].sum()), "mean rew", g["rew"].mean(), "wraps:",
          int((np.diff(g["cursor"][:, :, 0], axis=0) < 0).sum()))


def gen_w165_eval(out="w165_eval.npz", seed=0, n_episodes=3, n_steps=40):
    """MimicWalker165cm65kgEnv in evaluation mode (mimic_env.py:245,536-537 -> base_ref_trajecs.py:70-77): every episode
    starts at sample 0 of the recording.  Same synthetic recording and reference loading as gen_w165_rollout."""
    import scipy.io as spio
    from drloco_b200.ref_trajecs.loco3d_trajecs import synthetic_loco3d, N_ROWS
    proj = tempfile.mkdtemp()
    os.makedirs(os.path.join(proj, "mocaps/loco3d"))
    ang, vel = synthetic_loco3d()
    spio.savemat(os.path.join(proj, "mocaps/loco3d/loco3d_guoping.mat"),
                 {"angJoi": ang, "angDJoi": vel, "rowNameIK": np.array([f"row{i}" for i in range(N_ROWS)], dtype=object)})
    Env, Monitor, utils = load_reference_w165(proj)
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    e = Env()
    pristine = e.refs._qpos_full.copy()
    e.activate_evaluation()
    nv, nu, D = 19, 13, 47
    E, T = n_episodes, n_steps
    g = dict(actions=np.zeros((E, T, nu), np.float32), obs=np.full((E, T, D), np.nan), rew=np.full((E, T), np.nan),
             done=np.zeros((E, T), np.uint8), qpos=np.full((E, T, nv), np.nan), cursor=np.full((E, T, 2), -1, np.int32),
             obs0=np.zeros((E, D)), qpos0=np.zeros((E, nv)), cursor0=np.zeros((E, 2), np.int32),
             n_valid=np.zeros(E, np.int32))
    for k in range(E):
        e.refs._qpos_full[...] = pristine                        # Q4 waiver
        g["obs0"][k] = e.reset()
        g["qpos0"][k] = e.sim.data.qpos
        g["cursor0"][k] = (e.refs._pos, e.ep_dur)
        for t in range(T):
            a = (0.1 * (k + 1) * rng.uniform(-1, 1, nu)).astype(np.float32)
            g["actions"][k, t] = a
            o, r, d, _ = e.step(a)
            g["obs"][k, t], g["rew"][k, t], g["done"][k, t] = o, r, d
            g["qpos"][k, t] = e.sim.data.qpos
            g["cursor"][k, t] = (e.refs._pos, e.ep_dur)
            g["n_valid"][k] = t + 1
            if d:
                break
    g["meta"] = np.array("reference MimicWalker165cm65kgEnv (ENV_ID / mirroring set in the config) in evaluation mode over "
                         "oracle physics on synthetic_loco3d(seed 0); Q4 waived; seed=%d" % seed)
    np.savez_compressed(os.path.join(REPO, "tests/golden", out), **g)
    print(out, "valid steps per episode:", g["n_valid"], "first cursors:", g["cursor0"].tolist())


def gen_w3d_rollout(n_envs=8, n_steps=400, seed=0, out="w3d_rollout.npz", hypers_subst=()):
    Env, Monitor, utils = load_reference(hypers_subst)
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    envs, ewa, pristine = [], [], []
    for i in range(n_envs):
        utils._exp_weighted_averages = {}
        e = Monitor(Env())
        envs.append(e)
        ewa.append({})
        pristine.append(copy.deepcopy(e.env.refs.data))
    nv, nu, D = 14, 8, 29

    rsi_log = []

    # record the RSI draw of every reset
    for i, mon in enumerate(envs):
        refs = mon.env.refs
        orig = refs.get_random_init_state

        def wrapped(orig=orig, refs=refs, i=i):
            out = orig()
            rsi_log.append((i, refs._i_step, refs._pos))
            return out
        refs.get_random_init_state = wrapped

    T = n_steps
    g = dict(actions=np.zeros((T, n_envs, nu), np.float32), obs=np.zeros((T, n_envs, D)), rew=np.zeros((T, n_envs)),
             done=np.zeros((T, n_envs), np.uint8), terminal_obs=np.full((T, n_envs, D), np.nan),
             qpos=np.zeros((T, n_envs, nv)), qvel=np.zeros((T, n_envs, nv)), cursor=np.zeros((T, n_envs, 4), np.int32),
             ctrl=np.zeros((T, n_envs, nu)), comps=np.zeros((T, n_envs, 3)), walked=np.zeros((T, n_envs)),
             des_vel=np.zeros((T, n_envs)), rsi=np.full((T + 1, n_envs, 2), -1, np.int32),
             et=np.zeros((T, n_envs, 4), np.uint8))
    obs0 = np.zeros((n_envs, D))
    random.seed(seed + 1)
    for i in range(n_envs):
        utils._exp_weighted_averages = ewa[i]
        envs[i].env.refs.data = copy.deepcopy(pristine[i])
        obs0[i] = envs[i].env.reset()
        g["rsi"][0, i] = rsi_log[-1][1:]
    g["obs0"] = obs0
    g["qpos0"] = np.stack([e.env.sim.data.qpos.copy() for e in envs])
    g["qvel0"] = np.stack([e.env.sim.data.qvel.copy() for e in envs])
    g["cursor0"] = np.array([[e.env.refs._i_step, e.env.refs._pos, e.env.refs.count_steps_same_vel, e.env.ep_dur]
                             for e in envs], np.int32)
    for t in range(T):
        a = rng.uniform(-1.3, 1.3, size=(n_envs, nu)).astype(np.float32)
        # damp the action noise so that some envs survive long enough to cross mocap steps
        a[: n_envs // 2] *= 0.15
        g["actions"][t] = a
        for i, mon in enumerate(envs):
            utils._exp_weighted_averages = ewa[i]
            e = mon.env
            o, r, d, _ = mon.step(a[i])
            g["qpos"][t, i], g["qvel"][t, i] = e.sim.data.qpos, e.sim.data.qvel
            g["cursor"][t, i] = (e.refs._i_step, e.refs._pos, e.refs.count_steps_same_vel, e.ep_dur)
            g["ctrl"][t, i] = e.sim.data.ctrl
            g["comps"][t, i] = (e.pos_rew, e.vel_rew, e.com_rew)
            g["walked"][t, i] = e.walked_distance
            g["des_vel"][t, i] = e.desired_walking_speed[0]
            g["rew"][t, i], g["done"][t, i] = r, d
            g["et"][t, i] = e.do_terminate_early()        # dead code in step() (mimic_env.py:122-123), called here
            if d:
                g["terminal_obs"][t, i] = o
                e.refs.data = copy.deepcopy(pristine[i])  # Q4 waiver
                o = e.reset()
                g["rsi"][t + 1, i] = rsi_log[-1][1:]
            g["obs"][t, i] = o
    # Monitor attributes the callback reads through get_attr (callback.py:106-108,142,162-164,227)
    for name in ("ep_len_smoothed", "ep_ret_smoothed", "mean_reward_smoothed", "moved_distance",
                 "mean_ep_pos_rew_smoothed", "mean_ep_vel_rew_smoothed", "mean_ep_com_rew_smoothed",
                 "mean_abs_ep_torque_smoothed", "median_abs_torque_smoothed"):
        g["mon_" + name] = np.array([float(getattr(m, name)) for m in envs])
    g["mon_ep_lens"] = np.array([len(m.ep_lens) for m in envs], np.int32)
    g["mon_ep_lens_flat"] = np.array([x for m in envs for x in m.ep_lens], np.int32)
    # per-episode position records (monitor_wrapper.py:91-93,104-107,123-124); the construction-time Monitor state is
    # empty, so the lists hold exactly the episodes of this run
    g["mon_rsi_positions"] = np.array([x for m in envs for x in m.rsi_positions], np.int32)
    g["mon_et_positions"] = np.array([x for m in envs for x in m.et_positions], np.int32)
    g["mon_difficult_rsi_phases"] = np.array([x for m in envs for x in m.difficult_rsi_phases], np.int32)
    g["mon_n_rsi_per_env"] = np.array([len(m.rsi_positions) for m in envs], np.int32)
    g["meta"] = np.array("reference MimicWalker3dEnv+Monitor (unmodified) over oracle physics; Q4 waived; "
                         "per-env smoothing dicts; seed=%d" % seed)
    os.makedirs(os.path.join(REPO, "tests/golden"), exist_ok=True)
    np.savez_compressed(os.path.join(REPO, "tests/golden", out), **g)
    print(out, "episodes:", int(g["done"].sum()), "mean rew", g["rew"].mean())


def gen_w3d_eval(out="w3d_eval.npz", seed=0, n_episodes=6, n_steps=70):
    """evaluation mode (MimicEnv.activate_evaluation, mimic_env.py:245, :536-537): deterministic initial states
    (straight:237-265).  Each episode runs n_steps control steps with small actions (long enough to cross the first
    mocap-step transition), then the next reset is forced; a done ends it earlier."""
    Env, Monitor, utils = load_reference()
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    e = Env()
    pristine = copy.deepcopy(e.refs.data)
    e.activate_evaluation()
    nv, nu, D = 14, 8, 29
    E, T = n_episodes, n_steps
    g = dict(actions=np.zeros((E, T, nu), np.float32), obs=np.full((E, T, D), np.nan), rew=np.full((E, T), np.nan),
             done=np.zeros((E, T), np.uint8), qpos=np.full((E, T, nv), np.nan), cursor=np.full((E, T, 4), -1, np.int32),
             obs0=np.zeros((E, D)), qpos0=np.zeros((E, nv)), cursor0=np.zeros((E, 4), np.int32),
             n_valid=np.zeros(E, np.int32), phase=np.full((E, T), np.nan), left=np.zeros((E, T), np.uint8))
    count0 = e.refs.count_steps_same_vel
    for k in range(E):
        e.refs.data = copy.deepcopy(pristine)                   # Q4 waiver
        g["obs0"][k] = e.reset()
        g["qpos0"][k] = e.sim.data.qpos
        g["cursor0"][k] = (e.refs._i_step, e.refs._pos, e.refs.count_steps_same_vel, e.ep_dur)
        for t in range(T):
            a = (0.1 * rng.uniform(-1, 1, nu)).astype(np.float32)
            g["actions"][k, t] = a
            o, r, d, _ = e.step(a)
            g["obs"][k, t], g["rew"][k, t], g["done"][k, t] = o, r, d
            g["qpos"][k, t] = e.sim.data.qpos
            g["cursor"][k, t] = (e.refs._i_step, e.refs._pos, e.refs.count_steps_same_vel, e.ep_dur)
            g["phase"][k, t] = e.refs.get_phase_variable()
            g["left"][k, t] = e.refs.is_step_left()
            g["n_valid"][k] = t + 1
            if d:
                break
    g["count_at_construction"] = np.int32(count0)
    g["meta"] = np.array("reference MimicWalker3dEnv (unmodified) in evaluation mode over oracle physics; Q4 waived; seed=%d"
                         % seed)
    np.savez_compressed(os.path.join(REPO, "tests/golden", out), **g)
    print(out, "valid steps per episode:", g["n_valid"], "first cursors:", g["cursor0"].tolist())


def gen_w3d_blowup(out="w3d_blowup.npz", n_envs=2, n_steps=30):
    """the MujocoException path (mimic_env.py:82-91): the env resets itself, returns (obs, 0, True, {}) and the VecEnv
    resets it once more (Q19).  A blow-up is provoked by writing |qvel| > 1e10 into the simulator before chosen steps
    (MuJoCo's mj_checkVel limit; the fake sim raises MujocoException exactly then).  RSI draws are not logged here:
    generator and test seed Python's `random` identically before every env step / reset and make the same calls."""
    Env, Monitor, utils = load_reference()
    rng = np.random.default_rng(5)
    envs, ewa, pristine = [], [], []
    for i in range(n_envs):
        utils._exp_weighted_averages = {}
        e = Monitor(Env())
        envs.append(e)
        ewa.append({})
        pristine.append(copy.deepcopy(e.env.refs.data))
    nv, nu, D = 14, 8, 29
    T = n_steps
    blow = {(5, 0), (17, 1), (18, 1)}                       # (step, env): also two blow-ups in a row
    g = dict(actions=np.zeros((T, n_envs, nu), np.float32), obs=np.zeros((T, n_envs, D)), rew=np.zeros((T, n_envs)),
             done=np.zeros((T, n_envs), np.uint8), terminal_obs=np.full((T, n_envs, D), np.nan),
             qpos=np.zeros((T, n_envs, nv)), cursor=np.zeros((T, n_envs, 4), np.int32),
             blow=np.array(sorted(blow), np.int32), obs0=np.zeros((n_envs, D)),
             count0=np.array([e.env.refs.count_steps_same_vel for e in envs], np.int32))
    for i in range(n_envs):
        utils._exp_weighted_averages = ewa[i]
        envs[i].env.refs.data = copy.deepcopy(pristine[i])
        random.seed(900 + i)
        g["obs0"][i] = envs[i].env.reset()
    for t in range(T):
        a = (0.2 * rng.uniform(-1, 1, size=(n_envs, nu))).astype(np.float32)
        g["actions"][t] = a
        for i, mon in enumerate(envs):
            utils._exp_weighted_averages = ewa[i]
            e = mon.env
            if (t, i) in blow:
                e.sim.data.qvel[3] = 1e11
                e.refs.data = copy.deepcopy(pristine[i])     # Q4 waiver for the reset inside step()
            random.seed(1000 * t + i)
            o, r, d, _ = mon.step(a[i])
            g["rew"][t, i], g["done"][t, i] = r, d
            if d:
                g["terminal_obs"][t, i] = o
                e.refs.data = copy.deepcopy(pristine[i])
                random.seed(500000 + 1000 * t + i)
                o = e.reset()
            g["obs"][t, i] = o
            g["qpos"][t, i] = e.sim.data.qpos
            g["cursor"][t, i] = (e.refs._i_step, e.refs._pos, e.refs.count_steps_same_vel, e.ep_dur)
    for name in ("ep_len_smoothed", "ep_ret_smoothed", "mean_reward_smoothed", "moved_distance",
                 "mean_ep_pos_rew_smoothed", "mean_abs_ep_torque_smoothed", "median_abs_torque_smoothed"):
        g["mon_" + name] = np.array([float(getattr(m, name)) for m in envs])
    g["mon_ep_lens_flat"] = np.array([x for m in envs for x in m.ep_lens], np.int32)
    g["mon_et_positions"] = np.array([x for m in envs for x in m.et_positions], np.int32)
    np.savez_compressed(os.path.join(REPO, "tests/golden", out), **g)
    print(out, "dones at", np.argwhere(g["done"]).tolist(), "rewards there", g["rew"][g["done"] > 0])


def gen_w3d_cursor(out="w3d_cursor.npz", seed=0, n=1200):
    """pure cursor trace (SURVEY.md §8c known-answer iii): refs.next() from random.seed(0)."""
    _, _, _ = load_reference()
    from drloco.ref_trajecs import straight_walk_trajecs as ref_sw
    from drloco.mujoco.mimic_walker3d import qpos_indices, qvel_indices
    random.seed(seed)
    refs = ref_sw.StraightWalkingTrajectories(qpos_indices, qvel_indices)
    refs.get_random_init_state()
    tr = np.zeros((n + 1, 4), np.int32)
    ph = np.zeros(n + 1)
    dv = np.zeros(n + 1)
    qp = np.zeros((n + 1, 14))
    qv = np.zeros((n + 1, 14))
    left = np.zeros(n + 1, np.uint8)
    for t in range(n + 1):
        tr[t] = (refs._i_step, refs._pos, refs._trajec_len, refs.count_steps_same_vel)
        ph[t], dv[t] = refs.get_phase_variable(), refs.get_step_velocity()
        qp[t], qv[t] = refs.get_qpos().astype(np.float64), refs.get_qvel().astype(np.float64)
        left[t] = refs.is_step_left()
        refs.next()
    np.savez_compressed(os.path.join(REPO, "tests/golden", out), trace=tr, phase=ph, des_vel=dv, qpos=qp, qvel=qv,
                        left=left, step_velocities=np.asarray(refs.step_velocities, np.float64),
                        left_step_indices=np.asarray(refs.left_step_indices, np.int32))
    print(out, tr[0], tr[-1])


def gen_ref_config(out="ref_config.json"):
    """the step-path and learner settings of the reference's own config modules (drloco/config/config.py, hypers.py,
    drloco/mujoco/config.py), read from the imported modules: pins EnvConfig / PPOConfig defaults."""
    import json
    load_reference()
    from drloco.config import config as cfgl
    from drloco.config import hypers as h
    from drloco.mujoco import config as mjc
    g = dict(ENV_ID=cfgl.ENV_ID, CTRL_FREQ=cfgl.CTRL_FREQ, EVAL_N_TIMES=cfgl.EVAL_N_TIMES,
             MIN_STABLE_DISTANCE=cfgl.MIN_STABLE_DISTANCE, sim_freqs=dict(mjc.sim_freqs),
             modification=h.modification, mirr_py=bool(h.is_mod(h.MOD_MIRR_POLICY)),
             rew_weights=str(h.rew_weights), rew_scale=h.rew_scale, alive_bonus=h.alive_bonus, ep_dur_max=h.ep_dur_max,
             gamma=h.gamma, init_logstd=h.init_logstd, minibatch_size=h.minibatch_size, batch_size=h.batch_size,
             lr_start=h.lr_start, lr_final=h.lr_final, lr_scale=h.lr_scale, mio_samples=h.mio_samples, n_envs=h.n_envs,
             hid_layer_sizes=list(h.hid_layer_sizes), activation_fns=[f.__name__ for f in h.activation_fns],
             cliprange=h.cliprange, ent_coef=h.ent_coef, noptepochs=h.noptepochs,
             meta="values of the reference's config modules as imported under the stubs (is_remote() true)")
    json.dump(g, open(os.path.join(REPO, "tests/golden", out), "w"), indent=1)
    print(out, g)


SPEED_SPLAT = ("*self.desired_walking_speed,", "*np.atleast_1d(self.desired_walking_speed),")


def gen_w3d_speed_control(out="w3d_speed_control.npz", seed=0, n_episodes=3, n_steps=70):
    """MimicEnv.activate_speed_control (mimic_env.py:298-327): the generated profile, and rollouts with the profile
    driving the desired-velocity observation (:406-408) from deterministic initial states (:536-537).

    As shipped the path cannot produce an observation: `_get_obs` splats the profile's scalar entry
    (`*self.desired_walking_speed`, :429) and raises TypeError.  `mode == "as_shipped"` records exactly that (message in
    the fixture); `mode == "patched"` (own process) loads mimic_env.py with the one token wrapped in np.atleast_1d - the
    evident intent, a 1-vector like the mocap branch returns - and records the rollouts."""
    mode = os.environ.get("SPEED_MODE", "patched")
    if mode == "as_shipped":
        Env, Monitor, utils = load_reference()
        e = Env()
        e.activate_speed_control([0.5, 1.0, 0.75], 4)
        try:
            e.reset()
            msg = "no error"
        except TypeError as ex:
            msg = "TypeError: %s" % ex
        print("as shipped:", msg)
        return msg
    Env, Monitor, utils = load_reference(mimic_env_subst=[SPEED_SPLAT])
    random.seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    e = Env()
    pristine = copy.deepcopy(e.refs.data)
    count0 = e.refs.count_steps_same_vel
    g = {}
    cases = [([0.5, 1.0, 0.75], 4), ([1.0, 1.0], 10), ([0, 1, 2, 3], 1), ([0.6, 1.2, 0.9], 0.25)]
    for i, (speeds, dur) in enumerate(cases):
        e.activate_speed_control(speeds, dur)
        g["profile_%d" % i] = np.asarray(e.desired_walking_speed_trajectory, np.float64)
        g["profile_%d_args" % i] = np.array(list(speeds) + [dur], np.float64)
    # the last profile (50 control steps) stays active: the rollouts wrap around it (ep_dur % len)
    nv, nu, D = 14, 8, 29
    E, T = n_episodes, n_steps
    g.update(actions=np.zeros((E, T, nu), np.float32), obs=np.full((E, T, D), np.nan), rew=np.full((E, T), np.nan),
             done=np.zeros((E, T), np.uint8), qpos=np.full((E, T, nv), np.nan),
             cursor=np.full((E, T, 4), -1, np.int32), obs0=np.zeros((E, D)), qpos0=np.zeros((E, nv)),
             cursor0=np.zeros((E, 4), np.int32), n_valid=np.zeros(E, np.int32))
    for k in range(E):
        e.refs.data = copy.deepcopy(pristine)                   # Q4 waiver
        g["obs0"][k] = e.reset()
        g["qpos0"][k] = e.sim.data.qpos
        g["cursor0"][k] = (e.refs._i_step, e.refs._pos, e.refs.count_steps_same_vel, e.ep_dur)
        for t in range(T):
            a = (0.1 * rng.uniform(-1, 1, nu)).astype(np.float32)
            g["actions"][k, t] = a
            o, r, d, _ = e.step(a)
            g["obs"][k, t], g["rew"][k, t], g["done"][k, t] = o, r, d
            g["qpos"][k, t] = e.sim.data.qpos
            g["cursor"][k, t] = (e.refs._i_step, e.refs._pos, e.refs.count_steps_same_vel, e.ep_dur)
            g["n_valid"][k] = t + 1
            if d:
                break
    g["count_at_construction"] = np.int32(count0)
    g["as_shipped_error"] = np.array(os.environ.get("SPEED_AS_SHIPPED", ""))
    g["meta"] = np.array("reference MimicWalker3dEnv over oracle physics with activate_speed_control; mimic_env.py loaded "
                         "with %r -> %r (as shipped the observation raises, see as_shipped_error); Q4 waived; seed=%d"
                         % (SPEED_SPLAT[0], SPEED_SPLAT[1], seed))
    np.savez_compressed(os.path.join(REPO, "tests/golden", out), **g)
    print(out, "valid steps per episode:", g["n_valid"], "profile lengths:",
          [len(g["profile_%d" % i]) for i in range(len(cases))])


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "speed":                     # separate processes: one module substitution in the second
        import subprocess
        env = dict(os.environ, SPEED_MODE="as_shipped")
        msg = subprocess.run([sys.executable, __file__, "speed_worker"], env=env, capture_output=True, text=True)
        line = [x for x in msg.stdout.splitlines() if x.startswith("as shipped:")][0][len("as shipped: "):]
        env = dict(os.environ, SPEED_MODE="patched", SPEED_AS_SHIPPED=line)
        subprocess.check_call([sys.executable, __file__, "speed_worker"], env=env)
    if which == "speed_worker":
        gen_w3d_speed_control()
    if which == "config":
        gen_ref_config()
    if which in ("all", "cursor"):
        gen_w3d_cursor()
    if which in ("all", "rollout"):
        gen_w3d_rollout()
    if which in ("all", "blowup"):
        gen_w3d_blowup()
    if which in ("all", "eval"):
        gen_w3d_eval()
    if which == "timeout":                   # separate process: ep_dur_max = 25 in the reference's hypers.py (:58)
        gen_w3d_rollout(n_envs=4, n_steps=120, seed=3, out="w3d_timeout.npz",
                        hypers_subst=[("ep_dur_max = 3000", "ep_dur_max = 25")])
    if which == "w165":                      # separate process: the reference's config module is per-ENV_ID
        gen_w165_rollout()
    if which == "w165_eval":
        gen_w165_eval()
