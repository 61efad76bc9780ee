"""The numpy env oracle (oracle/env_oracle.py) against golden vectors produced by the reference's own Python
(tools/gen_golden.py): cursor indices / done flags bit-exact, float64 values bit-exact or to 1e-12."""
import os

import numpy as np
import pytest

from drloco_b200.walkers import make_spec
from oracle.env_oracle import OracleVecEnv, RefCursor
from oracle.physics import OraclePhysics


@pytest.fixture(scope="module")
def spec():
    return make_spec()


def test_mocap_known_answers(spec):
    """constants of the constant-speed mocap."""
    t = spec.mocap
    assert t.n_steps == 30 and t.n_samples == 7906
    assert list(t.step_len[:5]) == [262, 269, 268, 268, 249] and int(t.step_len[-1]) == 275
    assert list(np.nonzero(t.left_step)[0]) == list(range(1, 30, 2))
    assert abs(t.step_vel[0] - 1.471227) < 1e-6 and abs(t.step_vel[29] - 1.446132) < 1e-6
    assert np.all(t.ref[t.step_off, 0] < 0.005)      # COM-X starts at 0 on every step


def test_cursor_trace_matches_reference(spec, golden_dir):
    g = np.load(os.path.join(golden_dir, "w3d_cursor.npz"))
    t = spec.mocap
    np.testing.assert_array_equal(np.nonzero(t.left_step)[0], g["left_step_indices"])
    np.testing.assert_array_equal(t.step_vel, g["step_velocities"])
    c = RefCursor(t, 14)
    c.init_random(int(g["trace"][0, 0]), int(g["trace"][0, 1]))
    for k in range(g["trace"].shape[0]):
        assert (c.i_step, c.pos, c.len, c.count_steps_same_vel) == tuple(int(x) for x in g["trace"][k]), k
        assert c.get_phase_variable() == g["phase"][k]
        assert c.get_desired_walking_velocity_vector()[0] == g["des_vel"][k]
        np.testing.assert_array_equal(c.get_qpos(), g["qpos"][k])
        np.testing.assert_array_equal(c.get_qvel(), g["qvel"][k])
        assert c.is_step_left() == bool(g["left"][k])
        c.next()
    # known-answer (iii): transitions after RSI (27,197)
    tr = g["trace"]
    change = [k for k in range(1, len(tr)) if tr[k, 0] != tr[k - 1, 0]][:5]
    # (trace row k = the cursor after k calls of next(); the survey counts env steps, i.e. one less: 40, 174, ...)
    assert change == [41, 175, 312, 443, 577]
    assert [tuple(int(x) for x in tr[k]) for k in change] == [(28, 1, 269, 2), (29, 1, 275, 3), (0, 1, 262, 3),
                                                              (1, 2, 269, 4), (2, 2, 268, 5)]
    np.testing.assert_allclose(g["des_vel"][change], [1.464117, 1.464117, 1.471227, 1.471227, 1.471227], atol=5e-7)
    np.testing.assert_array_equal(g["phase"][change], [1 / 269, 1 / 275, 1 / 262, 2 / 269, 2 / 268])


@pytest.mark.parametrize("fixture,ep_dur_max", [("w3d_rollout.npz", 3000), ("w3d_timeout.npz", 25)])
def test_rollout_matches_reference(golden_dir, fixture, ep_dur_max):
    """reference MimicWalker3dEnv + Monitor over the oracle physics vs OracleVecEnv on the same actions and RSI draws.
    w3d_timeout.npz: the reference run with hypers.ep_dur_max = 25, every episode ends by time-out (reward +0.0, Q1)."""
    from drloco_b200.config import EnvConfig
    spec = make_spec(EnvConfig(ep_dur_max=ep_dur_max))
    g = np.load(os.path.join(golden_dir, fixture))
    if ep_dur_max == 25:
        assert g["done"].sum() >= 10 and (g["mon_ep_lens_flat"] == 25).all() and not np.signbit(g["rew"][g["done"] > 0]).any()
    T, N = g["actions"].shape[:2]
    venv = OracleVecEnv(spec, N, lambda: OraclePhysics(spec.model))
    for e in venv.envs:                                    # the reference env has stepped once in __init__ (Q14):
        e.env.refs.count_steps_same_vel = 1                # cursor count is all that survives into the first reset
    obs = venv.reset(g["rsi"][0, :, 0], g["rsi"][0, :, 1])
    # count_steps_same_vel persists across the construction-time step and reset (Q3): take it from the fixture
    for i, e in enumerate(venv.envs):
        e.env.refs.count_steps_same_vel = int(g["cursor0"][i, 2])
    np.testing.assert_array_equal(obs, g["obs0"])
    np.testing.assert_array_equal(np.stack([e.env.qpos for e in venv.envs]), g["qpos0"])
    for t in range(T):
        obs, rew, done, infos = venv.step(g["actions"][t], g["rsi"][t + 1, :, 0], g["rsi"][t + 1, :, 1])
        np.testing.assert_array_equal(done.astype(np.uint8), g["done"][t], err_msg=f"t={t}")
        np.testing.assert_array_equal(rew, g["rew"][t], err_msg=f"t={t}")
        assert np.array_equal(np.signbit(rew), np.signbit(g["rew"][t]))       # -0.0 on a fall (Q1)
        np.testing.assert_array_equal(obs, g["obs"][t], err_msg=f"t={t}")
        for i in range(N):
            if done[i]:
                np.testing.assert_array_equal(infos[i]["terminal_observation"], g["terminal_obs"][t, i])
        if "et" in g.files:        # do_terminate_early() of the reference env on the same states (mimic_env.py:652-702)
            np.testing.assert_array_equal([m.env.et_flags for m in venv.envs], g["et"][t].astype(bool), err_msg=f"t={t}")
    for name in ("ep_len_smoothed", "ep_ret_smoothed", "mean_reward_smoothed", "moved_distance",
                 "mean_ep_pos_rew_smoothed", "mean_ep_vel_rew_smoothed", "mean_ep_com_rew_smoothed",
                 "mean_abs_ep_torque_smoothed", "median_abs_torque_smoothed"):
        got = np.array([float(getattr(m, name)) for m in venv.envs])
        np.testing.assert_allclose(got, g["mon_" + name], rtol=1e-12, atol=0, err_msg=name)
    np.testing.assert_array_equal([x for m in venv.envs for x in m.ep_lens], g["mon_ep_lens_flat"])
    if "mon_rsi_positions" in g.files:
        for name in ("rsi_positions", "et_positions", "difficult_rsi_phases"):
            np.testing.assert_array_equal([x for m in venv.envs for x in getattr(m, name)], g["mon_" + name], err_msg=name)


def test_w165_rollout_matches_reference(golden_dir):
    """reference MimicWalker165cm65kgEnv + Monitor (wrap cursor base:95-103, joint-phase estimates mimic_env.py:330-360,
    2-D desired velocity loco3d:51-68, no mirroring) over the oracle physics vs OracleVecEnv, same actions / RSI draws.
    The recording is synthetic_loco3d(seed 0) on both sides (the reference read it from a .mat with its own schema)."""
    from drloco_b200.config import EnvConfig
    g = np.load(os.path.join(golden_dir, "w165_rollout.npz"))
    spec = make_spec(EnvConfig(env_id="MimicWalker165cm65kg"))
    T, N = g["actions"].shape[:2]
    assert spec.obs_dim == g["obs"].shape[2] == 47 and spec.act_dim == 13
    venv = OracleVecEnv(spec, N, lambda: OraclePhysics(spec.model))
    zeros = np.zeros(N, np.int32)
    obs = venv.reset(zeros, g["rsi"][0])
    np.testing.assert_array_equal(obs, g["obs0"])
    np.testing.assert_array_equal(np.stack([e.env.qpos for e in venv.envs]), g["qpos0"])
    np.testing.assert_array_equal([[e.env.refs.pos, e.env.ep_dur] for e in venv.envs], g["cursor0"])
    wraps = 0
    for t in range(T):
        before = np.array([e.env.refs.pos for e in venv.envs])
        obs, rew, done, infos = venv.step(g["actions"][t], zeros, g["rsi"][t + 1])
        np.testing.assert_array_equal(done.astype(np.uint8), g["done"][t], err_msg=f"t={t}")
        np.testing.assert_array_equal(rew, g["rew"][t], err_msg=f"t={t}")
        assert np.array_equal(np.signbit(rew), np.signbit(g["rew"][t]))
        np.testing.assert_array_equal(obs, g["obs"][t], err_msg=f"t={t}")
        for i in range(N):
            if done[i]:
                np.testing.assert_array_equal(infos[i]["terminal_observation"], g["terminal_obs"][t, i])
            else:
                assert (venv.envs[i].env.refs.pos, venv.envs[i].env.ep_dur) == tuple(g["cursor"][t, i])
                wraps += int(venv.envs[i].env.refs.pos < before[i])
    assert wraps >= 2 and g["done"].sum() >= 5             # the fixture crosses the end of the recording and has resets
    for name in ("ep_len_smoothed", "ep_ret_smoothed", "mean_reward_smoothed", "moved_distance",
                 "mean_ep_pos_rew_smoothed", "mean_ep_vel_rew_smoothed", "mean_ep_com_rew_smoothed",
                 "mean_abs_ep_torque_smoothed", "median_abs_torque_smoothed"):
        got = np.array([float(getattr(m, name)) for m in venv.envs])
        np.testing.assert_allclose(got, g["mon_" + name], rtol=1e-12, atol=0, err_msg=name)
    np.testing.assert_array_equal([x for m in venv.envs for x in m.ep_lens], g["mon_ep_lens_flat"])


def test_eval_mode_matches_reference(spec, golden_dir):
    """evaluation mode: deterministic init states straight:237-265 with the reference's table aliasing (Q27: data and
    length of mocap step 0 until the first transition, `_i_step` / mirroring / next-step choice of step n)."""
    from oracle.env_oracle import OracleMimicEnv
    g = np.load(os.path.join(golden_dir, "w3d_eval.npz"))
    env = OracleMimicEnv(spec, OraclePhysics(spec.model))
    env._EVAL_MODEL = True
    env.refs.count_steps_same_vel = int(g["count_at_construction"])    # after the construction-time step (Q14)
    E, T = g["actions"].shape[:2]
    crossed = 0
    for k in range(E):
        obs = env.reset()
        np.testing.assert_array_equal(obs, g["obs0"][k], err_msg=f"episode {k}")
        np.testing.assert_array_equal(env.qpos, g["qpos0"][k])
        assert (env.refs.i_step, env.refs.pos, env.refs.count_steps_same_vel, env.ep_dur) == tuple(g["cursor0"][k])
        for t in range(int(g["n_valid"][k])):
            before = env.refs.i_step
            obs, rew, done, _ = env.step(g["actions"][k, t])
            assert (env.refs.i_step, env.refs.pos, env.refs.count_steps_same_vel, env.ep_dur) == tuple(g["cursor"][k, t])
            assert env.refs.get_phase_variable() == g["phase"][k, t] and env.refs.is_step_left() == bool(g["left"][k, t])
            np.testing.assert_array_equal(obs, g["obs"][k, t], err_msg=f"episode {k} step {t}")
            np.testing.assert_array_equal(env.qpos, g["qpos"][k, t])
            assert rew == g["rew"][k, t] and done == bool(g["done"][k, t])
            crossed += int(env.refs.i_step != before)
    assert crossed >= E                                                 # every episode crossed its first transition
    # the quirk itself: episode 1 starts on step 1 (a left step: mirrored obs) but with step 0's phase denominator
    assert g["cursor0"][1, 0] == 1 and g["phase"][1, 0] == (g["cursor"][1, 0, 1]) / spec.mocap.step_len[0]


def test_w165_eval_mode_matches_reference(golden_dir):
    """MimicWalker165cm65kgEnv in evaluation mode: every episode starts at sample 0 of the recording
    (base_ref_trajecs.py:70-77 through mimic_env.py:536-537,575-579)."""
    from drloco_b200.config import EnvConfig
    from oracle.env_oracle import OracleMimicEnv
    g = np.load(os.path.join(golden_dir, "w165_eval.npz"))
    spec = make_spec(EnvConfig(env_id="MimicWalker165cm65kg"))
    env = OracleMimicEnv(spec, OraclePhysics(spec.model))
    env._EVAL_MODEL = True
    E = g["actions"].shape[0]
    for k in range(E):
        obs = env.reset()
        np.testing.assert_array_equal(obs, g["obs0"][k], err_msg=f"episode {k}")
        np.testing.assert_array_equal(env.qpos, g["qpos0"][k])
        assert (env.refs.pos, env.ep_dur) == tuple(g["cursor0"][k])
        for t in range(int(g["n_valid"][k])):
            obs, rew, done, _ = env.step(g["actions"][k, t])
            assert (env.refs.pos, env.ep_dur) == tuple(g["cursor"][k, t])
            np.testing.assert_array_equal(obs, g["obs"][k, t], err_msg=f"episode {k} step {t}")
            np.testing.assert_array_equal(env.qpos, g["qpos"][k, t])
            assert rew == g["rew"][k, t] and done == bool(g["done"][k, t])
    np.testing.assert_array_equal(g["cursor0"][:, 0], g["cursor0"][0, 0])        # the same start every time
    np.testing.assert_array_equal(g["obs0"][1], g["obs0"][0])


def test_speed_control_matches_reference(spec, golden_dir):
    """MimicEnv.activate_speed_control (mimic_env.py:298-327): profile generation, the desired-velocity observation
    driven by it (:406-408, index ep_dur % len) and the deterministic initial states it implies (:536-537).  As shipped the
    reference raises in _get_obs (scalar splat, :429) - recorded in the fixture; the rollouts come from the reference
    loaded with that one token wrapped in np.atleast_1d (tools/gen_golden.py::gen_w3d_speed_control)."""
    from drloco_b200.walkers import speed_profile
    from oracle.env_oracle import OracleMimicEnv
    g = np.load(os.path.join(golden_dir, "w3d_speed_control.npz"))
    assert str(g["as_shipped_error"]).startswith("TypeError")
    env = OracleMimicEnv(spec, OraclePhysics(spec.model))
    i = 0
    while "profile_%d" % i in g.files:
        args = g["profile_%d_args" % i]
        speeds, dur = list(args[:-1]), args[-1]
        dur = int(dur) if dur == int(dur) else float(dur)
        env.activate_speed_control(speeds, dur)
        np.testing.assert_array_equal(env.desired_walking_speed_trajectory, g["profile_%d" % i])
        np.testing.assert_array_equal(speed_profile(speeds, dur, spec.cfg.ctrl_freq), g["profile_%d" % i])   # host side of the GPU env
        i += 1
    assert i == 4 and len(env.desired_walking_speed_trajectory) == 50     # the last profile stays active
    env.refs.count_steps_same_vel = int(g["count_at_construction"])    # after the construction-time step (Q14)
    E = g["actions"].shape[0]
    wrapped = False
    for k in range(E):
        obs = env.reset()
        np.testing.assert_array_equal(obs, g["obs0"][k], err_msg=f"episode {k}")
        np.testing.assert_array_equal(env.qpos, g["qpos0"][k])
        assert (env.refs.i_step, env.refs.pos, env.refs.count_steps_same_vel, env.ep_dur) == tuple(g["cursor0"][k])
        for t in range(int(g["n_valid"][k])):
            obs, rew, done, _ = env.step(g["actions"][k, t])
            assert (env.refs.i_step, env.refs.pos, env.refs.count_steps_same_vel, env.ep_dur) == tuple(g["cursor"][k, t])
            np.testing.assert_array_equal(obs, g["obs"][k, t], err_msg=f"episode {k} step {t}")
            np.testing.assert_array_equal(env.qpos, g["qpos"][k, t])
            assert rew == g["rew"][k, t] and done == bool(g["done"][k, t])
            wrapped |= env.ep_dur > 50
    assert wrapped                                                       # the profile index wrapped around (ep_dur % len)
    # obs[1] is the profile entry at ep_dur before the step's increment (mimic_env.py:97,100; mirroring leaves slot 1 alone)
    np.testing.assert_array_equal(g["obs"][0, :60, 1], g["profile_3"][np.arange(60) % 50])


def test_blowup_path_matches_reference(spec, golden_dir):
    """MujocoException path (mimic_env.py:82-91, Q19): reset inside step(), reward 0, done, then the VecEnv's own
    reset.  Same `random` seeding protocol as tools/gen_golden.py (gen_w3d_blowup)."""
    import random
    from oracle.env_oracle import OracleMimicEnv, OracleMonitor
    g = np.load(os.path.join(golden_dir, "w3d_blowup.npz"))
    T, N = g["actions"].shape[:2]
    mons = [OracleMonitor(OracleMimicEnv(spec, OraclePhysics(spec.model))) for _ in range(N)]
    blow = {tuple(x) for x in g["blow"].tolist()}
    for i, m in enumerate(mons):
        m.env.refs.count_steps_same_vel = int(g["count0"][i])
        random.seed(900 + i)
        np.testing.assert_array_equal(m.env.reset(), g["obs0"][i])
    for t in range(T):
        for i, m in enumerate(mons):
            if (t, i) in blow:
                m.env.qvel[3] = 1e11
            random.seed(1000 * t + i)
            o, r, d, _ = m.step(g["actions"][t, i])
            assert d == bool(g["done"][t, i]) and r == g["rew"][t, i] and not np.signbit(r), (t, i)
            if d:
                np.testing.assert_array_equal(o, g["terminal_obs"][t, i])
                random.seed(500000 + 1000 * t + i)
                o = m.env.reset()
            np.testing.assert_array_equal(o, g["obs"][t, i], err_msg=f"t={t} env={i}")
            np.testing.assert_array_equal(m.env.qpos, g["qpos"][t, i])
            e = m.env
            assert (e.refs.i_step, e.refs.pos, e.refs.count_steps_same_vel, e.ep_dur) == tuple(g["cursor"][t, i])
    assert g["done"].sum() == 3 and (g["rew"][g["done"] > 0] == 0).all()
    np.testing.assert_array_equal([x for m in mons for x in m.ep_lens], g["mon_ep_lens_flat"])
    np.testing.assert_array_equal([x for m in mons for x in m.et_positions], g["mon_et_positions"])
    for name in ("ep_len_smoothed", "ep_ret_smoothed", "moved_distance", "mean_ep_pos_rew_smoothed",
                 "mean_abs_ep_torque_smoothed", "median_abs_torque_smoothed"):
        got = np.array([float(getattr(m, name)) for m in mons])
        np.testing.assert_allclose(got, g["mon_" + name], rtol=1e-12, err_msg=name)
    # env 1 had a one-step episode (two blow-ups in a row): the reference's np.mean of an empty list turns its
    # mean_reward_smoothed into NaN for good; oracle and kernel skip that update instead (waived)
    assert np.isnan(g["mon_mean_reward_smoothed"][1]) and np.isfinite(float(mons[1].mean_reward_smoothed))
    assert float(mons[0].mean_reward_smoothed) == g["mon_mean_reward_smoothed"][0]
