"""Host-side logic that needs no GPU: configuration semantics, mocap table compilers, mirror tables, sharding."""
import numpy as np
import pytest

from drloco_b200 import config as cfgm
from drloco_b200.ref_trajecs import (Loco3dReferenceTrajectories, StraightWalkingTrajectories, loco3d_trajecs as l3,
                                     straight_walk_trajecs as sw)
from drloco_b200.ref_trajecs.base_ref_trajecs import CURSOR_STEPWISE, CURSOR_WRAP
from drloco_b200.sharding import merge_moments, shard_range
from drloco_b200.walkers import (MIRROR_ACT_IDX, MIRROR_OBS_IDX, W165_QPOS_INDICES, make_spec, w3d_qpos_indices,
                                 w3d_qvel_indices)


def test_config_defaults_follow_the_reference():
    c = cfgm.EnvConfig()
    assert (c.env_id, c.ctrl_freq, c.frame_skip) == ("StraightMimicWalker", 200, 5)       # config.py:18-21
    assert c.rew_weights == (0.8, 0.2, 0.0, 0.0) and c.alive_bonus == 0.2 and c.ep_dur_max == 3000   # hypers.py:48-58
    assert c.gamma == 0.995 and c.is_mod(cfgm.MOD_MIRR_POLICY)                             # hypers.py:20-29,68
    w = cfgm.EnvConfig(env_id=cfgm.WALKER_165)
    assert (w.ctrl_freq, w.frame_skip, w.gamma) == (100, 10, 0.99)
    assert not w.is_mod(cfgm.MOD_MIRR_POLICY)                                              # hypers.py:31-39
    with pytest.raises(AssertionError):
        cfgm.EnvConfig(ctrl_freq=300, gamma=0.99).frame_skip                                         # mimic_env.py:203-206


def test_index_maps():
    assert w3d_qpos_indices(38) == [0, 1, 2, 35, 36, 37, 8, 7, 9, 10, 12, 11, 13, 14]
    assert w3d_qpos_indices(40)[3:6] == [37, 38, 39]                                       # ramp file (Q8)
    assert w3d_qvel_indices() == [15, 16, 17, 18, 19, 20, 22, 21, 23, 24, 26, 25, 27, 28]
    assert W165_QPOS_INDICES == [3, 5, 4, 1, 0, 2, 21, 20, 22, 6, 7, 8, 9, 10, 13, 14, 15, 16, 17]
    assert sorted(MIRROR_OBS_IDX) == list(range(29)) and sorted(MIRROR_ACT_IDX) == list(range(8))


def test_straight_walking_tables():
    s = make_spec()
    t = s.mocap
    assert t.cursor_mode == CURSOR_STEPWISE and t.increment == 2 and t.com_z_col == 2
    assert t.ref.shape == (7906, 28) and t.step_off[1] == t.step_len[0]
    assert (s.obs_dim, s.act_dim, s.frame_skip, s.mirror) == (29, 8, 5, True)
    oi, osn, ai, asn = s.mirror_tables()
    assert list(np.nonzero(osn < 0)[0]) == [2, 4, 6, 8, 12, 16, 18, 20, 22, 26]            # mimic_env.py:463
    assert list(np.nonzero(asn < 0)[0]) == [1, 5]                                          # mimic_env.py:486
    # mirroring twice is the identity
    x = np.arange(29, dtype=np.float64) + 1
    y = x[oi] * osn
    np.testing.assert_array_equal(y[oi] * osn, x)


def test_synthetic_mocaps_have_the_reference_schema(tmp_path):
    rows, lens = sw.synthetic_straight_rows(n_steps=6, seed=1)
    assert rows.shape[0] == 38 and rows.shape[1] == lens.sum()
    p = tmp_path / "syn.npz"
    np.savez(p, rows=rows, step_len=lens, sample_freq=400.0)
    t = StraightWalkingTrajectories(w3d_qpos_indices(38), w3d_qvel_indices(), path=str(p)).tables()
    assert t.n_steps == 6 and list(t.left_step) == [0, 1, 0, 1, 0, 1]
    ang, vel = l3.synthetic_loco3d(duration_s=4.0)
    assert ang.shape == vel.shape == (37, 2000)
    t = Loco3dReferenceTrajectories(W165_QPOS_INDICES, W165_QPOS_INDICES, {}).tables()
    assert t.cursor_mode == CURSOR_WRAP and t.increment == 5 and t.des_vel_window == 250 and t.n_steps == 1
    assert t.des_vel_prefix.shape == (t.n_samples + 1, 2)
    w = make_spec(cfgm.EnvConfig(env_id=cfgm.WALKER_165))
    assert (w.obs_dim, w.act_dim, w.frame_skip, w.mirror) == (47, 13, 10, False)


def test_adaptations_scale_rows():
    base = Loco3dReferenceTrajectories(W165_QPOS_INDICES, W165_QPOS_INDICES, {}).tables().ref
    t = Loco3dReferenceTrajectories(W165_QPOS_INDICES, W165_QPOS_INDICES, {l3.KNEE_ANG_R: 0.5}).tables().ref
    col = W165_QPOS_INDICES.index(l3.KNEE_ANG_R)
    np.testing.assert_allclose(t[:, col], 0.5 * base[:, col])
    np.testing.assert_allclose(t[:, 19 + col], 0.5 * base[:, 19 + col])                    # base:105-118 scales both


def test_shard_range_partitions_everything():
    for total, world in ((65536, 8), (4096, 3), (7, 2)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_merge_moments_is_running_mean_std():
    from oracle.env_oracle import RunningMeanStd
    rng = np.random.default_rng(0)
    rms = RunningMeanStd(shape=(5,))
    mean, var, count = np.zeros(5), np.ones(5), 1e-4
    for _ in range(4):
        x = rng.standard_normal((64, 5)) * 3 + 1
        rms.update(x)
        mean, var, count = merge_moments(mean, var, count, x.sum(0), (x * x).sum(0), float(len(x)))
        np.testing.assert_allclose(mean, rms.mean, rtol=1e-12)
        np.testing.assert_allclose(var, rms.var, rtol=1e-10)
        assert count == pytest.approx(rms.count)


def test_speed_profile_and_oracle_speed_control():
    """mimic_env.py:313-322 (profile), :406-408 (obs), :536-537 (deterministic init while following a profile)."""
    from drloco_b200.walkers import speed_profile
    from oracle.env_oracle import OracleMimicEnv
    from oracle.physics import OraclePhysics
    p = speed_profile([0.5, 1.0, 0.75], 4, 200)              # the reference docstring's example
    assert p.shape == (800,) and p[0] == 0.5 and p[399] == 1.0 and p[400] == 1.0 and p[-1] == 0.75
    assert np.allclose(np.diff(p[:400]), 0.5 / 399)
    assert speed_profile([1.0, 1.0], 10, 200).shape == (2000,)          # defaults of activate_speed_control
    assert speed_profile([0, 1, 2, 3], 1, 100).shape == (99,)           # int(100 / 3) = 33 per region
    spec = make_spec(cfgm.EnvConfig())
    env = OracleMimicEnv(spec, OraclePhysics(spec.model))
    env.activate_speed_control([0.5, 1.0], 1)
    obs = env.reset()
    assert (env.refs.i_step, env.ep_dur) == (0, 0) and obs[1] == 0.5    # deterministic init: step 0 at 75 %
    obs = env.step(np.zeros(8, np.float32))[0]
    assert obs[1] == env.desired_walking_speed_trajectory[0]            # _get_obs runs before ep_dur += 1
    obs = env.step(np.zeros(8, np.float32))[0]
    assert obs[1] == env.desired_walking_speed_trajectory[1]


def test_lazy_infos_behaves_like_a_list_of_dicts():
    """infos of VecEnv.step (SB3 VecEnv contract: one dict per env, 'terminal_observation' for finished ones)."""
    from drloco_b200.vec_env import LazyInfos
    t = np.arange(3, dtype=np.float32)
    infos = LazyInfos(5, {3: {"terminal_observation": t}})
    assert len(infos) == 5 and infos[0] == {} and infos[-2]["terminal_observation"] is t
    assert infos[0] is infos[0] and infos[0] is not infos[1]          # remembered once touched, never shared
    infos[1]["episode"] = {"r": 1.0}
    assert [sorted(d) for d in infos] == [[], ["episode"], [], ["terminal_observation"], []]
    assert infos[1:3] == [{"episode": {"r": 1.0}}, {}]
    assert infos == [{}, {"episode": {"r": 1.0}}, {}, {"terminal_observation": t}, {}]
    with pytest.raises(IndexError):
        infos[5]
    with pytest.raises(IndexError):
        infos[-6]
    assert "3" in repr(infos)


def test_terminal_record_buffer_is_parsed_into_infos():
    """layout written by csrc/vecnorm.cu::vecnorm_terminal_compact_kernel: 4 header words (count first), then one record
    { env index as int32 bits, d floats } per finished environment, in arbitrary order."""
    import torch
    from drloco_b200.vec_env import _TerminalRows
    n, d = 6, 3
    words = torch.zeros(_TerminalRows.words_for(n, d))
    tr = _TerminalRows(n, d, "cpu", host_words=words)
    assert tr.direct and tr.words == 4 + n * (d + 1)
    assert tr.collect(np.float32) == {}
    wi = words.numpy().view(np.int32)
    wf = words.numpy()
    wi[0] = 2
    wi[4], wf[5:8] = 4, [1.0, 2.0, 3.0]
    wi[8], wf[9:12] = 1, [-1.0, -2.0, -3.0]
    got = tr.collect(np.float32)
    assert sorted(got) == [1, 4]
    np.testing.assert_array_equal(got[4]["terminal_observation"], [1.0, 2.0, 3.0])
    np.testing.assert_array_equal(got[1]["terminal_observation"], [-1.0, -2.0, -3.0])
    wf[9] = 99.0                                   # the rows handed out are copies, not views of the reused buffer
    assert got[1]["terminal_observation"][0] == -1.0


def test_defaults_equal_the_reference_config_modules():
    """tests/golden/ref_config.json holds the values of the reference's own config modules as imported
    (tools/gen_golden.py config): EnvConfig / PPOConfig defaults must be those."""
    import json
    import os
    from drloco_b200.ppo import PPOConfig
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_config.json")))
    c, p = cfgm.EnvConfig(), PPOConfig()
    assert c.env_id == g["ENV_ID"] and c.ctrl_freq == g["CTRL_FREQ"] and cfgm.SIM_FREQS == g["sim_freqs"]
    assert c.eval_n_times == g["EVAL_N_TIMES"] and c.min_stable_distance == g["MIN_STABLE_DISTANCE"]
    assert "/".join(c.modifications) == g["modification"] and c.is_mod(cfgm.MOD_MIRR_POLICY) == g["mirr_py"]
    assert list(c.rew_weights) == json.loads(g["rew_weights"]) and c.rew_scale == g["rew_scale"]
    assert c.alive_bonus == g["alive_bonus"] and c.ep_dur_max == g["ep_dur_max"] and c.gamma == g["gamma"]
    assert p.gamma == g["gamma"] and p.init_logstd == g["init_logstd"] and p.minibatch_size == g["minibatch_size"]
    assert p.batch_size == g["batch_size"] and p.lr_start == g["lr_start"] and p.lr_final == g["lr_final"]
    assert g["lr_scale"] == 1                                              # plain linear decay lr_start -> lr_final
    assert p.total_steps == g["mio_samples"] * 10 ** 6 and list(p.hidden) == g["hid_layer_sizes"]
    assert g["activation_fns"] == ["Tanh", "Tanh"]                         # ActorCritic's trunk
    assert p.clip_range == g["cliprange"] and p.ent_coef == g["ent_coef"] and p.n_epochs == g["noptepochs"]
