"""Ingestion of the reference's mocap file formats.  The real ramp and loco3d recordings are not
in the reference checkout (.MISSING_LARGE_BLOBS), so files with the reference's *schema* are written here and loaded
through the same entry points a user would point at the real files:

  * straight walking, 40-row ramp layout: `Data` = object array of per-step (40 x T_i) matrices; GRF rows 35-36, trunk
    euler rows 37-39 (drloco/ref_trajecs/straight_walk_trajecs.py:22-27,85-91);
  * loco3d: `angJoi`, `angDJoi` (37 x T), `rowNameIK` (drloco/ref_trajecs/loco3d_trajecs.py:39-46);
  * `adaptations` row scaling (drloco/ref_trajecs/base_ref_trajecs.py:105-118).
The GPU halves replay the ingested tables kinematically (playback mode): every step must earn the maximal reward.
"""
import numpy as np
import pytest
import scipy.io as spio

from drloco_b200.config import EnvConfig
from drloco_b200.ref_trajecs import loco3d_trajecs as l3
from drloco_b200.ref_trajecs import straight_walk_trajecs as sw
from drloco_b200.walkers import W165_QPOS_INDICES, make_spec, w3d_qpos_indices, w3d_qvel_indices

W3D, W165 = "StraightMimicWalker", "MimicWalker165cm65kg"


def _ramp_mat(path, n_steps=12, seed=4):
    """the constant-speed rows re-laid out as the 40-row ramp file: two GRF rows inserted at 35-36"""
    rows38, lens = sw.synthetic_straight_rows(n_steps=n_steps, seed=seed)
    off = np.concatenate([[0], np.cumsum(lens)])
    data = np.empty(n_steps, dtype=object)
    for i in range(n_steps):
        s = rows38[:, off[i]:off[i + 1]]
        grf = np.stack([700.0 + 50.0 * np.sin(np.linspace(0, np.pi, s.shape[1])), np.zeros(s.shape[1])])
        data[i] = np.concatenate([s[:35], grf, s[35:38]], axis=0)
        assert data[i].shape[0] == 40
    spio.savemat(path, {"Data": data})
    return rows38, lens


def test_ramp_layout_mat_is_ingested(tmp_path):
    path = str(tmp_path / "Trajecs_Ramp_Slow_400Hz_EulerTrunkAdded.mat")
    rows38, lens = _ramp_mat(path)
    rows, step_len = sw.load_steps(path)
    assert rows.shape[0] == 40 and list(step_len) == list(lens)
    assert sw.trunk_euler_rows(40) == (37, 38, 39) and w3d_qpos_indices(40)[3:6] == [37, 38, 39]
    spec = make_spec(EnvConfig(env_id=W3D), mocap_path=path)
    t = spec.mocap
    assert t.n_steps == len(lens) and t.n_samples == int(lens.sum()) and t.ref.shape[1] == 28
    # the trunk euler columns come from rows 37-39 of the 40-row file = rows 35-37 of the 38-row layout
    np.testing.assert_array_equal(t.ref[:, 3:6], rows38[35:38].T)
    np.testing.assert_array_equal(t.ref[:, 0], rows38[sw.COM_POSX])
    np.testing.assert_array_equal(t.ref[:, 14:], rows38[w3d_qvel_indices()].T)
    # same tables as the 38-row file holding the same motion
    npz = str(tmp_path / "const.npz")
    np.savez(npz, rows=rows38, step_len=lens)
    t38 = make_spec(EnvConfig(env_id=W3D), mocap_path=npz).mocap
    np.testing.assert_array_equal(t.ref, t38.ref)
    np.testing.assert_array_equal(t.left_step, t38.left_step)
    np.testing.assert_array_equal(t.step_vel, t38.step_vel)
    assert t.left_step[1] == 1 and t.left_step[0] == 0               # odd synthetic steps swing the left leg


def _loco3d_mat(path, seconds=8.0, seed=3):
    ang, vel = l3.synthetic_loco3d(duration_s=seconds, seed=seed)
    names = np.array([f"row_{i}" for i in range(l3.N_ROWS)], dtype=object)
    spio.savemat(path, {"angJoi": ang, "angDJoi": vel, "rowNameIK": names})
    return ang, vel


def test_loco3d_mat_is_ingested_and_adaptations_scale_rows(tmp_path):
    path = str(tmp_path / "loco3d_guoping.mat")
    ang, vel = _loco3d_mat(path)
    refs = l3.Loco3dReferenceTrajectories(W165_QPOS_INDICES, W165_QPOS_INDICES, {}, path=path)
    t = refs.tables()
    assert len(refs.get_kinematics_labels()) == l3.N_ROWS                  # rowNameIK (loco3d:43)
    assert t.n_steps == 1 and t.n_samples == ang.shape[1] and t.increment == 5 and t.des_vel_window == 250
    np.testing.assert_array_equal(t.ref[:, :19], ang[W165_QPOS_INDICES].T)
    np.testing.assert_array_equal(t.ref[:, 19:], vel[W165_QPOS_INDICES].T)
    # desired velocity = mean pelvis x / z velocity over the next 0.5 s (loco3d:58-68) from the prefix sums
    pos = 1000
    want = [vel[l3.PELVIS_TX, pos:pos + 250].mean(), vel[l3.PELVIS_TZ, pos:pos + 250].mean()]
    got = (t.des_vel_prefix[pos + 250] - t.des_vel_prefix[pos]) / 250
    np.testing.assert_allclose(got, want, rtol=1e-12)
    # adaptations: row scalars applied to positions and velocities (base:105-118)
    ad = {l3.KNEE_ANG_R: 0.5, l3.PELVIS_TY: 1.1}
    t2 = l3.Loco3dReferenceTrajectories(W165_QPOS_INDICES, W165_QPOS_INDICES, ad, path=path).tables()
    ik, iy = W165_QPOS_INDICES.index(l3.KNEE_ANG_R), W165_QPOS_INDICES.index(l3.PELVIS_TY)
    np.testing.assert_allclose(t2.ref[:, ik], 0.5 * t.ref[:, ik], rtol=1e-15)
    np.testing.assert_allclose(t2.ref[:, 19 + ik], 0.5 * t.ref[:, 19 + ik], rtol=1e-15)
    np.testing.assert_allclose(t2.ref[:, iy], 1.1 * t.ref[:, iy], rtol=1e-15)
    others = [c for c in range(38) if c not in (ik, iy, 19 + ik, 19 + iy)]
    np.testing.assert_array_equal(t2.ref[:, others], t.ref[:, others])
    # the reference's StraightWalkingTrajectories drops `adaptations` (Q11): so does the mirror class
    a = sw.StraightWalkingTrajectories(w3d_qpos_indices(), w3d_qvel_indices(), adaptations={sw.KNEE_ANG_R: 0.5})
    b = sw.StraightWalkingTrajectories(w3d_qpos_indices(), w3d_qvel_indices())
    np.testing.assert_array_equal(a.tables().ref, b.tables().ref)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ramp", "loco3d"])
def test_ingested_mocap_plays_back_on_the_gpu(tmp_path, kind):
    """end to end: file -> tables -> device -> kinematic playback (mimic_env.py:265-293) earns the maximal reward on every
    step and follows the table rows exactly"""
    from drloco_b200.vec_env import B200MimicVecEnv
    if kind == "ramp":
        path = str(tmp_path / "ramp.mat")
        _ramp_mat(path)
        env_id, n = W3D, 8
    else:
        path = str(tmp_path / "loco3d.mat")
        _loco3d_mat(path)
        env_id, n = W165, 4
    cfg = EnvConfig(env_id=env_id)
    env = B200MimicVecEnv(env_id, num_envs=n, cfg=cfg, mocap_path=path, seed=2)
    t = env.spec.mocap
    out = env.playback_ref_trajectories(60)
    want = cfg.rew_scale * sum(cfg.rew_weights[:3]) + cfg.alive_bonus
    assert np.abs(out["reward"] - want).max() < 1e-6 and not out["done"].any()
    nv = env.spec.model.nv
    _, _, cur = env.get_state()
    for i in range(n):
        row = t.ref[t.step_off[cur[i, 0]] + cur[i, 1]]
        # joint angles (not the COM columns, which carry the per-episode x / z adjustments) equal the table row
        np.testing.assert_allclose(out["qpos"][-1, i, 3:], row[3:nv].astype(np.float32), rtol=0, atol=1e-6)
        np.testing.assert_allclose(out["qvel"][-1, i], row[nv:].astype(np.float32), rtol=0, atol=1e-5)
    env.close()
