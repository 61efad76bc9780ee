"""GPU parity tests: the CUDA path (through the C-ABI, via drloco_b200.vec_env) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): trajectory indices / phase / done / reset decisions bit-exact; qpos,
qvel and rewards within 1e-4 relative over a short horizon (fp32 device vs float64 oracle), divergence reported beyond."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from drloco_b200 import cabi  # noqa: E402
from drloco_b200.config import EnvConfig  # noqa: E402
from drloco_b200.walkers import make_spec  # noqa: E402

W3D, W165 = "StraightMimicWalker", "MimicWalker165cm65kg"
REL_TOL = 1e-4          # stated tolerance: |x_gpu - x_oracle|_inf <= REL_TOL * max(1, |x_oracle|_inf) per env


def _env(env_id=W3D, n=64, integrator="rk4", seed=5, ep_dur_max=3000, **kw):
    from drloco_b200.vec_env import B200MimicVecEnv
    cfg = EnvConfig(env_id=env_id, integrator=integrator, ep_dur_max=ep_dur_max)
    return B200MimicVecEnv(env_id, num_envs=n, cfg=cfg, seed=seed, **kw)


def _oracle(spec, n, integrator="rk4", physics=None):
    from oracle.env_oracle import OracleVecEnv
    from oracle.physics import OraclePhysics
    integ = cabi.INTEGRATOR_RK4 if integrator == "rk4" else cabi.INTEGRATOR_EULER
    return OracleVecEnv(spec, n, physics or (lambda: OraclePhysics(spec.model, integ)))


def _oracle_probed(spec, n, integrator="rk4"):
    """oracle env whose physics also runs the contact-sensitivity probe (oracle/sensitivity.py); returns (env, probes)"""
    from oracle.env_oracle import OracleVecEnv
    from oracle.sensitivity import SensitivityProbe
    integ = cabi.INTEGRATOR_RK4 if integrator == "rk4" else cabi.INTEGRATOR_EULER
    probes = []

    def make():
        probes.append(SensitivityProbe(spec.model, integ, seed=len(probes)))
        return probes[-1]
    return OracleVecEnv(spec, n, make), probes


def _rsi(spec, n, rng):
    t = spec.mocap
    istep = rng.integers(0, t.n_steps, n).astype(np.int32)
    pos = np.array([rng.integers(0, t.step_len[i]) for i in istep], np.int32)
    return istep, pos


def _rel_rows(x, ref):
    scale = np.maximum(1.0, np.abs(ref).max(axis=-1, keepdims=True))
    return (np.abs(x - ref) / scale).max(axis=-1)


def _rel(x, ref):
    return float(_rel_rows(x, ref).max())


def _ora_state(ora):
    q = np.stack([e.env.qpos for e in ora.envs])
    v = np.stack([e.env.qvel for e in ora.envs])
    c = np.array([[e.env.refs.i_step, e.env.refs.pos, e.env.refs.count_steps_same_vel, e.env.ep_dur] for e in ora.envs])
    return q, v, c


# --------------------------------------------------------------------------------------------------------------------
# one dynamics evaluation: mass matrix, bias force, constrained acceleration
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env_id", [W3D, W165])
def test_forward_dynamics_pieces(env_id):
    from oracle.physics import OraclePhysics
    n = 48
    env = _env(env_id, n, integrator="euler")
    spec = env.spec
    env.debug_set(frame_skip_override=1, enable_dump=True)
    m, t = spec.model, spec.mocap
    nv, nu = m.nv, m.nu
    rng = np.random.default_rng(0)
    P = OraclePhysics(m)
    rows = rng.integers(0, t.n_samples, n)
    q, v = t.ref[rows, :nv].copy(), t.ref[rows, nv:2 * nv].copy()
    for i in range(n):
        q[i, 3:] += 0.1 * rng.standard_normal(nv - 3)
        q[i, 0] = rng.uniform(-1, 30)                          # far from the origin: O-relative kinematics must not care
        q[i, 2] -= P.site_xpos(q[i])[:, 2].min() + rng.uniform(-0.004, 0.004)
        v[i] += 0.3 * rng.standard_normal(nv)
    q[0, 8 if env_id == W3D else 12] = -0.03 if env_id == W3D else 0.03     # a knee beyond its limit
    env.reset()
    cur = np.zeros((n, 4), np.int32)
    cur[:, 2] = 1
    env.set_state(q, v, cur)
    a = rng.uniform(-1, 1, (n, nu)).astype(np.float32)
    env.step(a)
    d = env.debug_read()
    ncon_total = 0
    for i in range(n):
        ctrl = (a[i] * np.float32(300)).astype(np.float64)
        qq, vv = q[i].astype(np.float32).astype(np.float64), v[i].astype(np.float32).astype(np.float64)
        Mo, co = P.mass_matrix(qq), P.bias(qq, vv)
        ao, dg = P.forward(qq, vv, ctrl)
        Mg, cg, ag = d[i, 2:2 + nv, :nv], d[i, 0, :nv], d[i, 2 + nv, :nv]
        assert np.abs(Mg - Mo).max() <= 2e-6 * np.abs(Mo).max()
        assert np.abs(cg - co).max() <= 2e-6 * max(1.0, np.abs(co).max())
        assert np.abs(ag - ao).max() <= 5e-4 * max(1.0, np.abs(ao).max()), (i, dg.ncon)
        assert int(d[i, 4 + nv, :].sum()) == dg.ncon                      # same contact set
        ncon_total += dg.ncon
    assert ncon_total >= n // 2                                            # the sample does exercise contacts
    env.close()


# --------------------------------------------------------------------------------------------------------------------
# reset (RSI + ground shift) and short-horizon rollouts
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env_id,integrator,steps", [(W3D, "rk4", 8), (W3D, "euler", 8), (W165, "rk4", 6)])
def test_rollout_parity_short_horizon(env_id, integrator, steps):
    """Stated tolerance: REL_TOL for EVERY environment the oracle does not flag as contact-sensitive.  An environment is
    flagged when the float64 oracle itself, re-run from its state perturbed at float32 resolution, instantiates a
    different constraint set within a control step (oracle/sensitivity.py) - there the soft-contact force is
    discontinuous and no float32 trajectory can be expected to follow the float64 one.  The flagged fraction is
    reported and bounded."""
    n = 64
    env = _env(env_id, n, integrator)
    spec = env.spec
    ora, probes = _oracle_probed(spec, n, integrator)
    rng = np.random.default_rng(1)
    istep, pos = _rsi(spec, n, rng)
    og, oo = env.reset(inject=(istep, pos)), ora.reset(istep, pos)
    assert _rel(og, oo) < 2e-5
    qg, vg, cg = env.get_state()
    qo, vo, co = _ora_state(ora)
    np.testing.assert_array_equal(cg, co)                                  # cursor after RSI + next(): bit-exact
    assert np.abs(qg - qo).max() < 1e-5
    curve, worst_flagged = [], 0.0
    for k in range(steps):
        a = rng.uniform(-1, 1, (n, spec.act_dim)).astype(np.float32)
        og, rg, dg, _ = env.step(a, inject=(istep, pos))
        oo, ro, do, _ = ora.step(a, istep, pos)
        qg, vg, cg = env.get_state()
        qo, vo, co = _ora_state(ora)
        for i in np.nonzero(do & dg)[0]:
            probes[i].clear()                                              # both sides restart from the same RSI state
        flagged = np.array([p.flagged for p in probes])
        rows = np.maximum(np.maximum(_rel_rows(qg, qo), _rel_rows(vg, vo)), np.abs(rg - ro))
        ok = ~flagged
        np.testing.assert_array_equal(dg[ok], do[ok], err_msg=f"done flags, step {k}")
        np.testing.assert_array_equal(cg[ok], co[ok], err_msg=f"cursor, step {k}")
        curve.append((float(rows[ok].max()), float(np.median(rows)), int(flagged.sum())))
        assert rows[ok].max() < REL_TOL, (k, np.nonzero(rows >= REL_TOL)[0], np.nonzero(flagged)[0], rows.max())
        assert np.median(rows) < 0.2 * REL_TOL
        if flagged.any():
            worst_flagged = max(worst_flagged, float(rows[flagged].max()))
        # phase and desired velocity are pure table lookups: bit-exact against float32(oracle)
        if spec.phase_from_cursor:
            left = np.array([bool(spec.mocap.left_step[c[0]]) for c in co])
            np.testing.assert_array_equal(og[ok, 0].astype(np.float32), oo[ok, 0].astype(np.float32))
            np.testing.assert_array_equal(og[ok, 1].astype(np.float32), oo[ok, 1].astype(np.float32))
            assert left.any() and (~left).any()
    n_flag = int(np.array([p.flagged for p in probes]).sum())
    assert n_flag <= n // 5, f"{n_flag} of {n} environments flagged contact-sensitive: the criterion would be vacuous"
    print(f"{env_id}/{integrator}: per step (max rel err of non-flagged envs, median, #flagged):",
          ["%.1e/%.1e/%d" % c for c in curve], f"; worst flagged env {worst_flagged:.1e}")
    env.close()


def test_golden_rollout_from_reference_python(golden_dir):
    """the first steps of the fixture produced by the reference's own MimicWalker3dEnv (tools/gen_golden.py)."""
    g = np.load(os.path.join(golden_dir, "w3d_rollout.npz"))
    n = g["actions"].shape[1]
    env = _env(W3D, n)
    env.reset(inject=(g["rsi"][0, :, 0], g["rsi"][0, :, 1]))
    cur = g["cursor0"].copy()
    env.set_state(g["qpos0"], g["qvel0"], cur)                  # includes count_steps_same_vel after the construction step (Q14)
    for t in range(6):
        obs, rew, done, _ = env.step(g["actions"][t])
        qg, vg, cg = env.get_state()
        np.testing.assert_array_equal(done, g["done"][t].astype(bool))
        np.testing.assert_array_equal(cg, g["cursor"][t])
        assert _rel(qg, g["qpos"][t]) < REL_TOL and _rel(vg, g["qvel"][t]) < REL_TOL
        assert np.abs(rew - g["rew"][t]).max() < REL_TOL
        assert _rel(obs, g["obs"][t]) < REL_TOL
        ex = env.extras().cpu().numpy()
        assert np.abs(ex[:, :3] - g["comps"][t]).max() < REL_TOL          # pos / vel / com reward components
        assert np.abs(ex[:, 3] - g["walked"][t]).max() < 1e-5
    env.close()


def test_w165_golden_rollout_from_reference_python(golden_dir):
    """the first steps of the fixture produced by the reference's own MimicWalker165cm65kgEnv (tools/gen_golden.py w165):
    wrap cursor, joint-phase estimates, windowed 2-D desired velocity, whole-recording COM-Z shift."""
    g = np.load(os.path.join(golden_dir, "w165_rollout.npz"))
    n = g["actions"].shape[1]
    env = _env(W165, n)
    zeros = np.zeros(n, np.int32)
    obs = env.reset(inject=(zeros, g["rsi"][0]))
    assert _rel(obs, g["obs0"]) < 2e-5
    qg, vg, cg = env.get_state()
    assert _rel(qg, g["qpos0"]) < 1e-6
    np.testing.assert_array_equal(cg[:, [1, 3]], g["cursor0"])
    for t in range(5):
        obs, rew, done, infos = env.step(g["actions"][t], inject=(zeros, g["rsi"][t + 1].clip(min=0)))
        qg, vg, cg = env.get_state()
        np.testing.assert_array_equal(done, g["done"][t].astype(bool))
        live = ~done
        np.testing.assert_array_equal(cg[live][:, [1, 3]], g["cursor"][t][live])
        assert _rel(qg[live], g["qpos"][t][live]) < REL_TOL and _rel(vg[live], g["qvel"][t][live]) < REL_TOL
        assert np.abs(rew - g["rew"][t]).max() < REL_TOL
        assert _rel(obs, g["obs"][t]) < REL_TOL
        ex = env.extras().cpu().numpy()
        assert np.abs(ex[live, :3] - g["comps"][t][live]).max() < REL_TOL
    env.close()


def test_w165_eval_mode_golden_from_reference_python(golden_dir):
    """MimicWalker165cm65kgEnv in evaluation mode against the reference-generated fixture (tools/gen_golden.py
    w165_eval): every episode starts at sample 0 of the recording (base_ref_trajecs.py:70-77); cursor for whole
    episodes, states / observations / rewards over the first steps (the oracle's sensitivity probe flags none of them)."""
    g = np.load(os.path.join(golden_dir, "w165_eval.npz"))
    env = _env(W165, 1)
    env.env_method("activate_evaluation")
    for k in range(g["actions"].shape[0]):
        obs = env.reset()
        assert _rel(obs, g["obs0"][k][None]) < 2e-5
        qg, vg, cg = env.get_state()
        np.testing.assert_array_equal(cg[0, [1, 3]], g["cursor0"][k])
        assert _rel(qg, g["qpos0"][k][None]) < 1e-6
        for t in range(int(g["n_valid"][k])):
            obs, rew, done, _ = env.step(g["actions"][k, t][None])
            assert not done[0]
            qg, vg, cg = env.get_state()
            np.testing.assert_array_equal(cg[0, [1, 3]], g["cursor"][k, t])
            if t < 5:
                assert _rel(obs, g["obs"][k, t][None]) < REL_TOL and abs(rew[0] - g["rew"][k, t]) < REL_TOL
                assert _rel(qg, g["qpos"][k, t][None]) < REL_TOL
    env.close()


def test_eval_mode_golden_from_reference_python(golden_dir):
    """evaluation mode against the fixture produced by the reference's own env (tools/gen_golden.py eval): deterministic
    init states incl. the reference's table aliasing (Q27) - cursor, phase, desired speed and mirroring for whole
    episodes (they do not depend on the physics), states and rewards over the first steps of each."""
    g = np.load(os.path.join(golden_dir, "w3d_eval.npz"))
    env = _env(W3D, 1)
    env.env_method("activate_evaluation")
    E, T = g["actions"].shape[:2]
    for k in range(E):
        obs = env.reset()
        assert _rel(obs, g["obs0"][k][None]) < 2e-5
        qg, vg, cg = env.get_state()
        np.testing.assert_array_equal(cg[0], g["cursor0"][k])
        assert _rel(qg, g["qpos0"][k][None]) < 1e-6
        for t in range(int(g["n_valid"][k])):
            obs, rew, done, _ = env.step(g["actions"][k, t][None])
            assert not done[0]
            qg, vg, cg = env.get_state()
            np.testing.assert_array_equal(cg[0], g["cursor"][k, t])
            assert abs(obs[0, 0] - g["phase"][k, t]) < 1e-6 and abs(obs[0, 1] - g["obs"][k, t, 1]) < 1e-6
            if t < 6:
                assert _rel(obs, g["obs"][k, t][None]) < REL_TOL and abs(rew[0] - g["rew"][k, t]) < REL_TOL
                assert _rel(qg, g["qpos"][k, t][None]) < REL_TOL
    env.close()


def test_speed_control_golden_from_reference_python(golden_dir):
    """speed control against the fixture produced by the reference's own env (tools/gen_golden.py speed; the reference
    loaded with the one token without which `_get_obs` raises, see the fixture's meta): deterministic init states, cursor
    and the profile-driven desired-velocity observation for whole episodes, states and rewards over the first steps."""
    g = np.load(os.path.join(golden_dir, "w3d_speed_control.npz"))
    env = _env(W3D, 1)
    args = g["profile_3_args"]
    env.env_method("activate_speed_control", list(args[:-1]), float(args[-1]))
    np.testing.assert_allclose(env.desired_walking_speed_trajectory, g["profile_3"], rtol=1e-7)
    E = g["actions"].shape[0]
    for k in range(E):
        obs = env.reset()
        assert _rel(obs, g["obs0"][k][None]) < 2e-5
        qg, vg, cg = env.get_state()
        np.testing.assert_array_equal(cg[0], g["cursor0"][k])
        assert _rel(qg, g["qpos0"][k][None]) < 1e-6
        for t in range(int(g["n_valid"][k])):
            obs, rew, done, _ = env.step(g["actions"][k, t][None])
            assert not done[0]
            qg, vg, cg = env.get_state()
            np.testing.assert_array_equal(cg[0], g["cursor"][k, t])
            assert abs(obs[0, 1] - g["obs"][k, t, 1]) < 1e-6
            if t < 6:
                assert _rel(obs, g["obs"][k, t][None]) < REL_TOL and abs(rew[0] - g["rew"][k, t]) < REL_TOL
                assert _rel(qg, g["qpos"][k, t][None]) < REL_TOL
    env.close()


def test_long_horizon_divergence_is_reported():
    """beyond the short horizon fp32 and fp64 trajectories separate at contact events; termination decisions must
    still agree for as long as the states do."""
    n, steps = 64, 40
    env = _env(W3D, n)
    ora = _oracle(env.spec, n)
    rng = np.random.default_rng(2)
    istep, pos = _rsi(env.spec, n, rng)
    env.reset(inject=(istep, pos))
    ora.reset(istep, pos)
    med = []
    for k in range(steps):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        _, _, dg, _ = env.step(a, inject=(istep, pos))
        _, _, do, _ = ora.step(a, istep, pos)
        qg, _, _ = env.get_state()
        qo, _, _ = _ora_state(ora)
        err = np.abs(qg - qo).max(axis=1)
        med.append(float(np.median(err)))
        close = err < 1e-3
        np.testing.assert_array_equal(dg[close], do[close])
    print("median |q_gpu - q_oracle| per step:", ["%.1e" % x for x in med])
    assert med[9] < 1e-4 and med[-1] < 1e-2
    env.close()


# --------------------------------------------------------------------------------------------------------------------
# environment logic on injected states (frame_skip = 0): reward, termination, cursor, Monitor statistics
# --------------------------------------------------------------------------------------------------------------------
class _FrozenPhysics:
    """oracle-side stand-in for 'no physics': states are injected, site positions come from the real kinematics."""

    def __init__(self, model):
        from oracle.physics import OraclePhysics
        self._p = OraclePhysics(model)
        self.qacc_warm = self._p.qacc_warm

    def step(self, q, v, ctrl, nsub):
        return False

    def site_xpos(self, q):
        return self._p.site_xpos(q)


def test_env_logic_and_monitor_on_injected_states():
    n, steps = 32, 60
    from drloco_b200.vec_env import B200MimicVecEnv
    # short episodes: time-outs occur naturally (hypers.py:58); the torque history for Monitor's median statistic is on
    env = B200MimicVecEnv(W3D, num_envs=n, cfg=EnvConfig(env_id=W3D, ep_dur_max=25, median_torque=True), seed=5)
    spec = env.spec
    env.debug_set(frame_skip_override=0)
    ora = _oracle(spec, n, physics=lambda: _FrozenPhysics(spec.model))
    rng = np.random.default_rng(3)
    istep, pos = _rsi(spec, n, rng)
    pos = np.minimum(pos, spec.mocap.step_len[istep] - 1).astype(np.int32)
    pos[:4] = spec.mocap.step_len[istep[:4]] - 1                 # RSI on the last sample: immediate step transition
    og, oo = env.reset(inject=(istep, pos)), ora.reset(istep, pos)
    assert _rel(og, oo) < 2e-5
    n_done = 0
    for k in range(steps):
        # perturb the (frozen) states the same way on both sides; drop a few walkers below the fall height
        qg, vg, cg = env.get_state()
        dq = (0.02 * rng.standard_normal(qg.shape)).astype(np.float32)
        dv = (0.2 * rng.standard_normal(vg.shape)).astype(np.float32)
        q_new, v_new = qg + dq, vg + dv
        fall = rng.random(n) < 0.03
        q_new[fall, 2] = 0.45
        env.set_state(q_new, v_new, cg)
        for i, e in enumerate(ora.envs):
            e.env.qpos[:] = q_new[i].astype(np.float64)
            e.env.qvel[:] = v_new[i].astype(np.float64)
        a = rng.uniform(-1.5, 1.5, (n, 8)).astype(np.float32)
        og, rg, dg, ig = env.step(a, inject=(istep, pos))
        oo, ro, do, io = ora.step(a, istep, pos)
        np.testing.assert_array_equal(dg, do)
        np.testing.assert_array_equal(np.signbit(rg), np.signbit(ro))      # -0.0 on a fall, +0.0 on a time-out (Q1)
        assert np.abs(rg - ro).max() < 2e-5
        assert _rel(og, oo) < 2e-5
        for i in np.nonzero(do)[0]:
            assert _rel(ig[i]["terminal_observation"][None], io[i]["terminal_observation"][None]) < 2e-5
        _, _, cg2 = env.get_state()
        np.testing.assert_array_equal(cg2, _ora_state(ora)[2])
        n_done += int(do.sum())
    assert n_done > 20
    # Monitor attributes served through get_attr (callback.py:106-108,142,162-164)
    for name in ("ep_len_smoothed", "ep_ret_smoothed", "mean_reward_smoothed", "mean_ep_pos_rew_smoothed",
                 "mean_ep_vel_rew_smoothed", "mean_ep_com_rew_smoothed", "moved_distance",
                 "mean_abs_ep_torque_smoothed", "median_abs_torque_smoothed"):
        got = np.array(env.get_attr(name))
        want = np.array([float(getattr(m, name)) for m in ora.envs])
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5, err_msg=name)
    lens_o = sorted(x for m in ora.envs for x in m.ep_lens)
    assert sorted(env.episode_lengths().tolist()) == lens_o
    # Monitor's per-episode position records (monitor_wrapper.py:91-93,104-107,123-124); ring order is arbitrary within a step
    rec = env.episode_records()
    got = sorted(zip(rec["rsi_pos"].tolist(), rec["et_pos"].tolist(), rec["ep_len"].tolist(), rec["difficult"].tolist()))
    want = [t for m in ora.envs for t in zip(m.rsi_positions, m.et_positions, m.ep_lens, m.ep_difficult)]
    assert got == sorted(want)
    assert sorted(env.get_attr("et_positions")[0]) == sorted(x for m in ora.envs for x in m.et_positions)
    assert sorted(env.get_attr("difficult_rsi_phases")[0]) == sorted(x for m in ora.envs for x in m.difficult_rsi_phases)
    # rsi_positions also lists the episodes still running (monitor_wrapper.py:91-93)
    assert sorted(env.get_attr("rsi_positions")[0]) == sorted(x for m in ora.envs for x in m.rsi_positions)
    assert rec["difficult"].sum() >= 1
    st = env.stats()
    assert st["episodes"] == n_done and st["falls"] + st["timeouts"] == n_done and st["timeouts"] >= 3
    assert st["ep_len_sum"] == sum(lens_o)
    env.set_attr("ep_lens", [])                                  # callback.py:69-70
    assert len(env.episode_lengths()) == 0
    env.close()


def test_blowup_path_golden_from_reference_python(golden_dir):
    """MujocoException path (mimic_env.py:82-91) on the GPU against the fixture produced by the reference's own
    MimicWalker3dEnv + Monitor (tools/gen_golden.py gen_w3d_blowup): |qvel| > 1e10 written into the simulator before
    chosen steps, incl. two blow-ups in a row.  The RSI draws are those of the generator (same `random` seeding
    protocol as tests/test_oracle_golden.py::test_blowup_path_matches_reference), injected on the device.  Kept
    reference behaviour: reward +0.0, done, observation after the reset, cursor, Monitor episode records.  Waived
    (Q19): the reference resets twice and reports the first reset's observation as terminal."""
    import random
    g = np.load(os.path.join(golden_dir, "w3d_blowup.npz"))
    T, n = g["actions"].shape[:2]
    env = _env(W3D, n)
    t_ = env.spec.mocap

    def draw(seed):                                           # RefCursor.init_random / straight:460-474
        random.seed(seed)
        i = random.randint(0, t_.n_steps - 1)
        return i, random.randint(0, int(t_.step_len[i]) - 1)

    cur = np.zeros((n, 4), np.int32)
    cur[:, 2] = g["count0"]                                   # count_steps_same_vel survives resets (Q3)
    env.set_state(None, None, cur)
    rsi = np.array([draw(900 + i) for i in range(n)], np.int32)
    obs = env.reset(inject=(rsi[:, 0], rsi[:, 1]))
    assert _rel(obs, g["obs0"]) < 2e-5
    blow = {tuple(x) for x in g["blow"].tolist()}
    since_reset = np.zeros(n, int)
    for t in range(T):
        hit = [i for i in range(n) if (t, i) in blow]
        if hit:
            q, v, c = env.get_state()
            v[hit, 3] = 1e11
            env.set_state(q, v, c)
        rsi = np.array([draw(500000 + 1000 * t + i) for i in range(n)], np.int32)
        obs, rew, done, infos = env.step(g["actions"][t], inject=(rsi[:, 0], rsi[:, 1]))
        np.testing.assert_array_equal(done, g["done"][t].astype(bool), err_msg=f"t={t}")
        qg, _, cg = env.get_state()
        np.testing.assert_array_equal(cg, g["cursor"][t], err_msg=f"t={t}")
        since_reset = np.where(done, 0, since_reset + 1)
        for i in range(n):
            if done[i]:                                       # every done of this fixture is a blow-up
                assert (t, i) in blow and rew[i] == 0.0 and not np.signbit(rew[i])
                assert "terminal_observation" in infos[i] and np.isfinite(infos[i]["terminal_observation"]).all()
                assert _rel(obs[i][None], g["obs"][t, i][None]) < 2e-5 and _rel(qg[i][None], g["qpos"][t, i][None]) < 1e-5
            else:
                # free-running fp32 physics vs the float64 fixture: tight over the stated horizon, loose after it
                tol = REL_TOL if since_reset[i] <= 8 else 5e-3
                assert _rel(obs[i][None], g["obs"][t, i][None]) < tol and abs(rew[i] - g["rew"][t, i]) < tol, (t, i)
    assert env.stats()["blowups"] == 3 and env.stats()["episodes"] == 3
    rec = env.episode_records()
    assert sorted(rec["ep_len"].tolist()) == sorted(g["mon_ep_lens_flat"].tolist())
    # (Monitor.et_positions / moved_distance of a blown-up episode are read after the reference's first, in-step
    #  reset; with the single reset here they describe the state before it - part of the Q19 waiver)
    np.testing.assert_allclose(env.get_attr("ep_len_smoothed"), g["mon_ep_len_smoothed"], rtol=1e-6)
    env.close()


def test_blowup_path_and_eval_mode():
    n = 16
    env = _env(W3D, n)
    env.reset()
    q, v, c = env.get_state()
    v[3, 0] = 3e10 * 10                                          # MuJoCo would raise on |qvel| > 1e10 (mimic_env.py:86-91)
    q[5, 7] = np.nan
    env.set_state(q, v, c)
    obs, rew, done, infos = env.step(np.zeros((n, 8), np.float32))
    assert done[3] and done[5] and rew[3] == 0 and not np.signbit(rew[3])
    assert np.isfinite(obs).all() and "terminal_observation" in infos[3]
    assert env.stats()["blowups"] == 2
    # deterministic initialisation during evaluation (straight_walk_trajecs.py:237-265)
    env.env_method("activate_evaluation")
    for k in range(3):
        env.reset()
        _, _, c = env.get_state()
        L = env.spec.mocap.step_len[k]
        inc = env.spec.mocap.increment
        assert (c[:, 0] == k).all() and (c[:, 1] == (3 * L) // 4 + inc).all()
    # batched evaluation: env i plays the i-th consecutive evaluation episode of the reference (callback.py:296-297)
    env.set_det_init_counters(np.arange(n) % env.cfg.eval_n_times)
    env.reset()
    _, _, c = env.get_state()
    np.testing.assert_array_equal(c[:, 0], np.arange(n))
    np.testing.assert_array_equal(c[:, 1], (3 * env.spec.mocap.step_len[:n]) // 4 + env.spec.mocap.increment)
    with pytest.raises(Exception):
        env.set_det_init_counters(np.full(n, 99))
    env.close()


@pytest.mark.parametrize("enabled", [False, True])
def test_early_termination_check(enabled):
    """do_terminate_early (mimic_env.py:652-702): reasons counted always; ends the episode only when configured
    (the reference never lets it, mimic_env.py:120-123)."""
    from drloco_b200.vec_env import B200MimicVecEnv
    n, steps = 32, 40
    cfg = EnvConfig(env_id=W3D, early_termination=enabled)
    env = B200MimicVecEnv(W3D, num_envs=n, cfg=cfg, seed=5)
    spec = env.spec
    env.debug_set(frame_skip_override=0)
    ora = _oracle(spec, n, physics=lambda: _FrozenPhysics(spec.model))
    rng = np.random.default_rng(8)
    istep, pos = _rsi(spec, n, rng)
    env.reset(inject=(istep, pos)), ora.reset(istep, pos)
    env.reset_stats()
    want = np.zeros(3)
    n_done = 0
    for k in range(steps):
        qg, vg, cg = env.get_state()
        q_new = qg.copy()
        # push single quantities over their thresholds: COM-Y, COM-Z (between 0.5 and 0.75), frontal and sagittal trunk angle
        who = rng.integers(0, 6, n)
        q_new[who == 1, 1] = rng.choice([-0.25, 0.25])
        q_new[who == 2, 2] = 0.7
        q_new[who == 3, 3] += 0.3
        q_new[who == 4, 4] = rng.choice([-0.1, 0.35])
        env.set_state(q_new, vg, cg)
        for i, e in enumerate(ora.envs):
            e.env.qpos[:] = q_new[i].astype(np.float64)
            e.env.qvel[:] = vg[i].astype(np.float64)
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        og, rg, dg, _ = env.step(a, inject=(istep, pos))
        flags = []
        oo, ro, do = [], [], []
        for i, m in enumerate(ora.envs):                     # flags must be read before the auto-reset of OracleVecEnv
            o, r, d, _ = m.step(a[i])
            flags.append(m.env.et_flags)
            if d:
                o = m.env.reset(istep[i], pos[i])
            oo.append(o), ro.append(r), do.append(d)
        flags, do = np.array(flags), np.array(do)
        want += flags[:, 1:].sum(axis=0)
        np.testing.assert_array_equal(dg, do)
        assert _rel(og, np.stack(oo)) < 2e-5 and np.abs(rg - np.array(ro)).max() < 2e-5
        np.testing.assert_array_equal(np.signbit(rg), np.signbit(np.array(ro)))
        if enabled:
            np.testing.assert_array_equal(dg, flags[:, 0])   # nothing falls below 0.5 or times out in this test
        else:
            assert not dg.any()
        n_done += int(dg.sum())
    st = env.stats()
    assert [st["et_com_low"], st["et_trunk"], st["et_drunk"]] == want.tolist() and want.min() > 10
    assert st["falls"] == n_done and (n_done > 50 if enabled else n_done == 0)
    env.close()


def test_speed_control_profile():
    """env_method('activate_speed_control') (mimic_env.py:298-327,406-408,536-537): the desired-velocity observation
    follows the profile indexed by the episode duration, resets become deterministic."""
    n, steps = 16, 70
    env = _env(W3D, n, ep_dur_max=25)
    spec = env.spec
    env.debug_set(frame_skip_override=0)
    ora = _oracle(spec, n, physics=lambda: _FrozenPhysics(spec.model))
    speeds, dur = [0.5, 1.0, 0.75], 0.1                          # 2 regions x 10 control steps: shorter than an episode
    env.env_method("activate_speed_control", speeds, dur)
    for m in ora.envs:
        m.env.activate_speed_control(speeds, dur)
    prof = env.desired_walking_speed_trajectory
    assert len(prof) == 20 and prof[0] == 0.5 and prof[9] == 1.0 and prof[10] == 1.0 and prof[19] == 0.75
    og, oo = env.reset(), ora.reset()
    assert _rel(og, oo) < 2e-5
    np.testing.assert_array_equal(env.get_state()[2], _ora_state(ora)[2])       # deterministic init on both sides
    assert np.allclose(og[:, 1], prof[0])
    rng = np.random.default_rng(11)
    seen_wrap = False
    for k in range(steps):
        qg, vg, cg = env.get_state()
        q_new = qg + (0.02 * rng.standard_normal(qg.shape)).astype(np.float32)
        q_new[rng.random(n) < 0.03, 2] = 0.45
        env.set_state(q_new, vg, cg)
        for i, e in enumerate(ora.envs):
            e.env.qpos[:] = q_new[i].astype(np.float64)
            e.env.qvel[:] = vg[i].astype(np.float64)
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        ep_dur_before = cg[:, 3].copy()
        og, rg, dg, ig = env.step(a)
        oo, ro, do, io = ora.step(a)
        np.testing.assert_array_equal(dg, do)
        assert _rel(og, oo) < 2e-5 and np.abs(rg - ro).max() < 2e-5
        np.testing.assert_array_equal(env.get_state()[2], _ora_state(ora)[2])
        # obs[1] is the desired speed (the mirror keeps it in place): profile[ep_dur % len] at _get_obs time
        want = np.where(dg, prof[0], prof[ep_dur_before % len(prof)])
        np.testing.assert_allclose(og[:, 1], want, rtol=1e-6)
        seen_wrap = seen_wrap or bool(((ep_dur_before >= len(prof)) & ~dg).any())
    assert seen_wrap
    env.activate_speed_control([])                               # off again: mocap step velocity is back
    o = env.step(np.zeros((n, 8), np.float32))[0]
    assert np.all(np.abs(o[:, 1] - 1.45) < 0.1)
    env.close()


@pytest.mark.parametrize("env_id", [W3D, W165])
def test_playback_ref_trajectories(env_id):
    """kinematic playback (mimic_env.py:265-293): the state follows the mocap, so every step earns the maximal
    imitation reward; states and cursors equal the oracle's replay bit for bit (float32 copies of the same table)."""
    n, T = 8, 400
    env = _env(env_id, n)
    spec = env.spec
    ora = _oracle(spec, n, physics=lambda: _FrozenPhysics(spec.model))
    rng = np.random.default_rng(2)
    istep, pos = _rsi(spec, n, rng)
    env.set_playback(True)
    for m in ora.envs:
        m.env._PLAYBACK_REF_TRAJECS = True
    og, oo = env.reset(inject=(istep, pos)), ora.reset(istep, pos)
    assert _rel(og, oo) < 2e-5
    a = np.zeros((n, env.act_dim), np.float32)
    xs = []
    for k in range(T):
        og, rg, dg, _ = env.step(a, inject=(istep, pos))
        oo, ro, do, _ = ora.step(a, istep, pos)
        assert not dg.any() and not do.any()
        qg, vg, cg = env.get_state()
        qo, vo, co = _ora_state(ora)
        np.testing.assert_array_equal(cg, co)
        assert _rel(qg, qo) < 1e-6 and _rel(vg, vo) < 1e-6 and _rel(og, oo) < 2e-5
        want = env.cfg.rew_scale * sum(env.cfg.rew_weights[:3]) + env.cfg.alive_bonus
        assert np.abs(rg - want).max() < 1e-6 and np.abs(ro - want).max() < 1e-12
        xs.append(qg[:, 0].copy())
    if env_id == W3D:
        # several mocap steps were crossed; COM X is continuous over the first transition only, later ones jump back
        # because the travelled distance is always taken from the RSI step (Q2) - the oracle comparison above covers it
        assert (cg[:, 0] != istep).all()
        assert (np.diff(np.stack(xs), axis=0) < -0.3).any()
    env.set_playback(False)
    out = env.playback_ref_trajectories(20)                      # the reference-named entry point, renderer-free
    assert out["qpos"].shape == (20, n, spec.model.nv) and np.abs(out["reward"] - want).max() < 1e-6
    env.close()


# --------------------------------------------------------------------------------------------------------------------
# VecNormalize kernels
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("host_outputs", ["mapped", "copy"])
def test_vecnormalize_matches_sb3_semantics(host_outputs):
    from drloco_b200.vec_env import B200VecNormalize
    from oracle.env_oracle import RunningMeanStd
    n = 256 if host_outputs == "mapped" else 320          # 320 > _TerminalRows.PREFIX: the copy path's second fetch
    env = _env(W3D, n)
    vn = B200VecNormalize(env)
    vn.host_outputs = host_outputs
    D = env.obs_dim
    obs_rms, ret_rms, ret = RunningMeanStd(shape=(D,)), RunningMeanStd(shape=()), np.zeros(n)
    rng = np.random.default_rng(4)
    o = vn.reset()
    raw = env.obs.cpu().numpy().astype(np.float64)
    ret_rms.update(np.zeros(n))                    # SB3 1.0 VecNormalize.reset: ret = 0, ret_rms.update(ret); obs_rms untouched
    want = np.clip((raw - obs_rms.mean) / np.sqrt(obs_rms.var + 1e-8), -10, 10)
    assert np.abs(o - want).max() < 1e-4
    for k in range(12):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        o, r, d, _ = vn.step(a)
        raw, rr = env.obs.cpu().numpy().astype(np.float64), env.rew.cpu().numpy().astype(np.float64)
        obs_rms.update(raw)
        want_o = np.clip((raw - obs_rms.mean) / np.sqrt(obs_rms.var + 1e-8), -10, 10)
        ret = ret * 0.99 + rr
        ret_rms.update(ret)
        want_r = np.clip(rr / np.sqrt(ret_rms.var + 1e-8), -10, 10)
        ret[d] = 0
        assert np.abs(o - want_o).max() < 2e-4 and np.abs(r - want_r).max() < 2e-4
    np.testing.assert_allclose(vn.obs_rms.mean, obs_rms.mean, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(vn.obs_rms.var, obs_rms.var, rtol=1e-5)
    assert abs(vn.ret_rms.var - ret_rms.var) < 1e-5 * ret_rms.var and vn.obs_rms.count == pytest.approx(obs_rms.count)
    # terminal observations come back normalised with the current statistics (SB3 VecNormalize.step_wait)
    q, v, c = env.get_state()
    q[:5, 2] = 0.45
    env.set_state(q, v, c)
    if host_outputs == "copy":
        q[:, 2] = 0.45                           # more finished envs than the copied record prefix holds
        env.set_state(q, v, c)
    o, r, d, infos = vn.step(np.zeros((n, 8), np.float32))
    assert d[:5].all() and (host_outputs == "mapped" or d.all())
    raw_t = env.terminal_obs.cpu().numpy().astype(np.float64)
    m, var = vn.obs_rms.mean, vn.obs_rms.var
    for i in (range(5) if host_outputs == "mapped" else range(n)):
        want_t = np.clip((raw_t[i] - m) / np.sqrt(var + 1e-8), -10, 10)
        assert np.abs(infos[i]["terminal_observation"] - want_t).max() < 2e-4
    assert all("terminal_observation" not in infos[i] for i in np.nonzero(~d)[0])
    sd = vn.state_dict()
    vn2 = B200VecNormalize(env)
    vn2.load_state_dict(sd)
    np.testing.assert_array_equal(vn2.obs_rms.var, vn.obs_rms.var)
    env.close()


def test_statistics_are_bit_reproducible_and_stats_sync_every():
    """the moments / Monitor statistics come from fixed-order sums (no atomics): two runs of the same launch
    configuration agree bit for bit (another CTA shape groups the additions differently: same to rounding);
    stats_sync_every=K merges the accumulated moments of K steps at once (same statistics up to rounding at the merge
    steps, stale normalisation in between)."""
    from drloco_b200.vec_env import B200VecNormalize
    n, steps = 512, 12
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    acts = torch.rand(steps, n, 8, device="cuda", generator=g) * 2 - 1
    runs = []
    for block, every in ((128, 1), (128, 1), (64, 1), (128, 4)):
        env = _env(W3D, n, seed=21)
        env.debug_set(block_threads=block)
        vn = B200VecNormalize(env, stats_sync_every=every)
        vn.reset_tensor()
        outs = []
        for k in range(steps):
            o, r, d = vn.step_tensor(acts[k])
            outs.append((o.clone(), r.clone()))
        runs.append(dict(mean=vn.obs_rms.mean, var=vn.obs_rms.var, count=vn.obs_rms.count, rvar=vn.ret_rms.var,
                         outs=outs, stats=env.stats()))
        env.close()
    a, a2, b, c = runs
    np.testing.assert_array_equal(a["mean"], a2["mean"])
    np.testing.assert_array_equal(a["var"], a2["var"])
    assert a["rvar"] == a2["rvar"] and a["stats"] == a2["stats"]
    for (o1, r1), (o2, r2) in zip(a["outs"], a2["outs"]):
        assert torch.equal(o1, o2) and torch.equal(r1, r2)
    np.testing.assert_allclose(b["mean"], a["mean"], rtol=1e-12, atol=1e-15)      # other CTA shape: other grouping
    np.testing.assert_allclose(b["var"], a["var"], rtol=1e-10)
    assert b["stats"]["episodes"] == a["stats"]["episodes"] and b["stats"]["env_steps"] == n * steps
    # K = 4 over 12 steps: merged at steps 4, 8, 12 - the same data has been merged by the end
    assert c["count"] == a["count"]
    np.testing.assert_allclose(c["mean"], a["mean"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(c["var"], a["var"], rtol=1e-10)
    assert abs(c["rvar"] - a["rvar"]) <= 1e-10 * a["rvar"]
    assert not torch.equal(c["outs"][1][0], a["outs"][1][0])          # steps 1-3 were normalised with stale statistics


def test_reset_update_modes_seed_and_lazy_infos():
    from drloco_b200.vec_env import B200VecNormalize, LazyInfos
    n = 64
    env = _env(W3D, n)
    vn = B200VecNormalize(env, reset_update="obs")
    o = vn.reset()
    raw = env.obs.cpu().numpy().astype(np.float64)
    from oracle.env_oracle import RunningMeanStd
    rms = RunningMeanStd(shape=(env.obs_dim,))
    rms.update(raw)
    assert np.abs(o - np.clip((raw - rms.mean) / np.sqrt(rms.var + 1e-8), -10, 10)).max() < 1e-4
    assert vn.ret_rms.count == pytest.approx(1e-4)                     # "obs" mode leaves the return statistics alone
    # infos: list-like, fresh dict per entry, nothing shared between entries or steps
    _, _, d, infos = vn.step(np.zeros((n, 8), np.float32))
    assert isinstance(infos, LazyInfos) and len(infos) == n and infos[0] == {} and infos[0] is not infos[1]
    infos[0]["x"] = 1
    assert infos[0]["x"] == 1 and "x" not in infos[1]
    _, _, _, infos2 = vn.step(np.zeros((n, 8), np.float32))
    assert "x" not in infos2[0] and len(list(infos2)) == n
    # seed(): re-keys the RSI stream (utils.py:113): envs built with different seeds draw the same initial states once
    # they are given the same seed, and different ones otherwise
    env2 = _env(W3D, n, seed=99)
    assert env.seed(1234) == [1234 + i for i in range(n)] and env2.seed(1234)[0] == 1234
    env.debug_set(frame_skip_override=0)
    vn.reset()
    env2.reset()
    ca, cb = env.get_state()[2], env2.get_state()[2]
    assert not np.array_equal(ca[:, :2], cb[:, :2])                     # reset counters differ (env was reset before)
    env3 = _env(W3D, n, seed=7)
    env3.seed(1234)
    env3.reset()
    np.testing.assert_array_equal(env3.get_state()[2][:, :2], cb[:, :2])
    env3.seed(4321)
    env3.reset()
    env2.reset()
    assert not np.array_equal(env3.get_state()[2][:, :2], env2.get_state()[2][:, :2])
    for e in (env, env2, env3):
        e.close()


def test_vec_env_factory_save_load(tmp_path):
    """drop-in for drloco.common.utils.vec_env / save_model / load_env (utils.py:97-134,175-192,234-240)."""
    from drloco_b200.vec_env import vec_env
    env = vec_env(W3D, num_envs=32, seed=3, norm_rew=True)
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(5):
        env.step(rng.uniform(-1, 1, (32, 8)).astype(np.float32))
    path = str(tmp_path / "env_ckpt")
    env.save(path)
    env2 = vec_env(W3D, num_envs=1, seed=3, norm_rew=True, load_path=path)      # the callback's 1-env eval env
    np.testing.assert_array_equal(env2.obs_rms.mean, env.obs_rms.mean)
    np.testing.assert_array_equal(env2.obs_rms.var, env.obs_rms.var)
    assert env2.ret_rms.var == env.ret_rms.var and env2.num_envs == 1
    env2.training = False
    env2.env_method("activate_evaluation")
    o = env2.reset()
    assert o.shape == (1, 29) and np.isfinite(o).all() and np.abs(o).max() <= 10.0
    assert env.get_attr("ep_len_smoothed") == env.venv.get_attr("ep_len_smoothed")
    # the same statistics in the reference's own file format (pickled SB3 VecNormalize, utils.py:183-184) and back
    path_sb3 = str(tmp_path / "env_ckpt_sb3")
    env.save_sb3(path_sb3)
    env3 = vec_env(W3D, num_envs=2, seed=3, norm_rew=False, load_path=path_sb3)
    np.testing.assert_array_equal(env3.obs_rms.mean, env.obs_rms.mean)
    np.testing.assert_array_equal(env3.obs_rms.var, env.obs_rms.var)
    assert env3.obs_rms.count == env.obs_rms.count and env3.ret_rms.var == env.ret_rms.var
    env.close()
    env2.close()
    env3.close()


# --------------------------------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: 4096 envs) — size-independent invariants
# --------------------------------------------------------------------------------------------------------------------
def test_full_size_invariants_and_determinism():
    n, steps = 4096, 30
    outs = []
    for block in (128, 32):
        env = _env(W3D, n, seed=11)
        env.debug_set(block_threads=block)
        g = torch.Generator(device="cuda")
        g.manual_seed(0)
        acts = torch.rand(steps, n, 8, device="cuda", generator=g) * 3 - 1.5
        env.reset_tensor()
        tot_done = 0
        for k in range(steps):
            obs, rew, done = env.step_tensor(acts[k])
            tot_done += int(done.sum())
        q, v, c = env.get_state()
        o, r = obs.cpu().numpy(), rew.cpu().numpy()
        assert np.isfinite(o).all() and np.isfinite(q).all() and np.isfinite(v).all()
        assert (r >= 0).all() and (r <= 1.2 + 1e-6).all()                 # 0.8 + 0.2 + alive bonus 0.2
        t = env.spec.mocap
        assert (c[:, 0] >= 0).all() and (c[:, 0] < t.n_steps).all() and (c[:, 1] < t.step_len[c[:, 0]]).all()
        assert (o[:, 0] >= 0).all() and (o[:, 0] <= 1).all()              # phase variable (straight:173-177)
        assert (q[:, 2] >= 0.5 - 1e-6).all()                              # fallen walkers were reset (mimic_env.py:120)
        st = env.stats()
        assert st["env_steps"] == n * steps and st["episodes"] == tot_done
        outs.append((o, r, q, c))
        env.close()
    # same seed, different CTA shape: bitwise identical (no atomics or races on the state path)
    for x, y in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(x, y)


def test_c_abi_call_order_errors():
    import ctypes as C
    from drloco_b200 import lib
    l = lib.load()
    spec = make_spec()
    cfg = cabi.DrlConfig()
    cfg.num_envs, cfg.device, cfg.frame_skip, cfg.obs_dim, cfg.act_dim, cfg.ctrl_freq = 8, 0, 5, 29, 8, 200.0
    h = C.c_void_p()
    assert l.drl_create(C.byref(cfg), C.byref(h)) == 0
    dummy = torch.zeros(8 * 29, device="cuda")
    p = C.c_void_p(dummy.data_ptr())
    assert l.drl_step(h, p, p, p, p, None, None, None, None) == -3          # DRL_ERR_STATE: nothing uploaded
    assert b"uploaded" in l.drl_last_error()
    cm = cabi.pack_model(spec.model)
    cm.nv = 15
    assert l.drl_upload_model(h, C.byref(cm)) == -4                          # DRL_ERR_UNSUPPORTED
    assert l.drl_destroy(h) == 0
    # a batch whose state cannot be allocated (2^31-1 envs x 256 B of state rows > HBM): a clean error, nothing leaked,
    # and the library keeps working afterwards
    free0 = torch.cuda.mem_get_info()[0]
    cfg.num_envs = 2 ** 31 - 1
    h2 = C.c_void_p()
    assert l.drl_create(C.byref(cfg), C.byref(h2)) == 0
    cm = cabi.pack_model(spec.model)
    assert l.drl_upload_model(h2, C.byref(cm)) == -2                         # DRL_ERR_CUDA
    assert b"memory" in l.drl_last_error().lower()
    assert l.drl_step(h2, p, p, p, p, None, None, None, None) == -3          # still "nothing uploaded"
    assert l.drl_destroy(h2) == 0
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)
    env = _env(W3D, 8)
    env.reset()
    env.close()


def test_odd_env_count_and_masked_reset():
    """N = 7: the last warp carries one live and one padding environment; masked reset touches only the chosen envs."""
    n = 7
    env = _env(W3D, n)
    ora = _oracle(env.spec, n)
    rng = np.random.default_rng(6)
    istep, pos = _rsi(env.spec, n, rng)
    og, oo = env.reset(inject=(istep, pos)), ora.reset(istep, pos)
    assert _rel(og, oo) < 2e-5
    for k in range(4):
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        og, rg, dg, _ = env.step(a, inject=(istep, pos))
        oo, ro, do, _ = ora.step(a, istep, pos)
        np.testing.assert_array_equal(dg, do)
        assert _rel(og, oo) < REL_TOL and np.abs(rg - ro).max() < REL_TOL
    q0, v0, c0 = env.get_state()
    mask = torch.tensor([0, 1, 0, 0, 1, 0, 1], dtype=torch.uint8, device="cuda")
    env.reset_tensor(mask, inject=(istep, pos))
    q1, v1, c1 = env.get_state()
    keep = ~mask.cpu().numpy().astype(bool)
    np.testing.assert_array_equal(q1[keep], q0[keep])
    np.testing.assert_array_equal(c1[keep], c0[keep])
    assert (c1[~keep, 3] == 0).all() and (np.abs(q1[~keep] - q0[~keep]).max(axis=1) > 0).all()
    env.close()


def test_w165_config4_invariants():
    """BASELINE.json configs[3]: MimicWalker165cm65kg on the (synthetic) loco3d mocap, RSI + termination, 16384 envs."""
    n, steps = 16384, 12
    env = _env(W165, n, seed=3)
    assert (env.obs_dim, env.act_dim) == (47, 13) and env.launch_info()["lanes_per_env"] == 32
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    env.reset_tensor()
    done_total = 0
    for k in range(steps):
        obs, rew, done = env.step_tensor(torch.rand(n, 13, device="cuda", generator=g) * 2 - 1)
        done_total += int(done.sum())
    o, r = obs.cpu().numpy(), rew.cpu().numpy()
    q, v, c = env.get_state()
    assert np.isfinite(o).all() and np.isfinite(q).all() and (r >= 0).all() and (r <= 1.2 + 1e-6).all()
    assert (c[:, 0] == 0).all() and (c[:, 1] >= 0).all() and (c[:, 1] < env.spec.mocap.step_len[0] - 1).all()
    assert (c[:, 1] % env.spec.mocap.increment == 0).sum() >= 0          # base:95-103 advances by 5 samples
    assert (np.abs(o[:, 0:8:2]) <= 1 + 1e-6).all()                         # phase angles / pi (mimic_env.py:350-352)
    st = env.stats()
    assert st["env_steps"] == n * steps and st["episodes"] == done_total
    env.close()


def test_ppo_consumer_runs_on_device_rollouts(tmp_path):
    """next-tier smoke: the PPO loop (drloco_b200/ppo.py) consumes the tensor API end to end and produces finite updates."""
    from drloco_b200.ppo import PPO, PPOConfig, evaluate_walking
    from drloco_b200.vec_env import vec_env
    env = vec_env(W3D, num_envs=256, seed=1)
    cfg = PPOConfig(batch_size=256 * 16, minibatch_size=1024, total_steps=256 * 16 * 3)
    from drloco_b200.training_monitor import TrainingMonitor
    agent = PPO(env, cfg, seed=0)
    mon = TrainingMonitor(agent, env.venv.cfg, str(tmp_path) + "/")     # callback.py: evaluation + logging cadence
    mon.on_training_start()
    agent.step_callback = mon.on_step
    agent.learn(log_every=1)
    mon.on_training_end()
    assert agent.num_timesteps == 256 * 16 * 3 and len(agent.log) == 3
    assert mon.num_timesteps == agent.num_timesteps                      # one on_step per env.step
    assert len(mon.moved_distances) == 10 and mon.min_episode_duration >= 1   # evaluated once (10 episodes < 1M steps)
    assert mon.count_stable_walks == 0 and os.listdir(str(tmp_path) + "/models") == []   # untrained: checkpoint deleted
    agent.save(str(tmp_path / "m.zip"))
    w0 = agent.policy.action_net.weight.clone()
    agent.policy.action_net.weight.data.zero_()
    agent.load(str(tmp_path / "m.zip"))
    assert torch.equal(agent.policy.action_net.weight, w0)
    for row in agent.log:
        assert all(np.isfinite(v) for v in row.values())
        assert 0.0 < row["mean_step_reward"] <= 1.2
    ev = evaluate_walking(agent.policy, env, n_episodes=4)
    assert ev["n_episodes"] == 4 and ev["min_episode_duration"] >= 1 and np.isfinite(ev["mean_walked_distance"])
    assert len(ev["moved_distances"]) == len(ev["ep_durs"]) == len(ev["mean_rewards"]) == 4
    env.close()
