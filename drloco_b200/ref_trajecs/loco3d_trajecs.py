"""loco3d mocap (one continuous 500 Hz recording, 37 IK rows) -> device tables.

Mirror of reference drloco/ref_trajecs/loco3d_trajecs.py.  The reference's data file
(mocaps/loco3d/loco3d_guoping.mat, fields angJoi / angDJoi / rowNameIK) is not shipped with the reference checkout
(.MISSING_LARGE_BLOBS), so ``synthetic_loco3d`` generates a recording with the same schema.
"""
from __future__ import annotations

import numpy as np

from .base_ref_trajecs import BaseReferenceTrajectories, MocapTables, CURSOR_WRAP, gather

# row-index constants (loco3d:7-19)
PELVIS_TILT, PELVIS_LIST, PELVIS_ROTATION = range(0, 3)
PELVIS_TX, PELVIS_TY, PELVIS_TZ = range(3, 6)
HIP_FLEXION_R, HIP_ADDUCTION_R, HIP_ROTATION_R = range(6, 9)
KNEE_ANG_R, ANKLE_ANG_R = range(9, 11)
SUBTALAR_ANG_R, MTP_ANG_R = range(11, 13)
HIP_FLEXION_L, HIP_ADDUCTION_L, HIP_ROTATION_L = range(13, 16)
KNEE_ANG_L, ANKLE_ANG_L = range(16, 18)
SUBTALAR_ANG_L, MTP_ANG_L = range(18, 20)
LUMBAR_EXTENSION, LUMBAR_BENDING, LUMBAR_ROTATION = range(20, 23)
ARM_FLEX_R, ARM_ADD_R, ARM_ROT_R = range(23, 26)
ELBOW_FLEX_R, PRO_SUP_R, WRIST_FLEX_R, WRIST_DEV_R = range(26, 30)
ARM_FLEX_L, ARM_ADD_L, ARM_ROT_L = range(30, 33)
ELBOW_FLEX_L, PRO_SUP_L, WRIST_FLEX_L, WRIST_DEV_L = range(33, 37)
N_ROWS = 37


def synthetic_loco3d(duration_s=60.0, sample_freq=500.0, seed=0):
    """(angJoi, angDJoi) float64 [37, T]: periodic gait with slow speed / heading modulation."""
    rng = np.random.default_rng(seed)
    T = int(duration_s * sample_freq)
    t = np.arange(T) / sample_freq
    f = 0.9                                   # stride frequency [Hz]
    ph = 2 * np.pi * f * t
    speed = 1.2 + 0.25 * np.sin(2 * np.pi * t / 17.0)
    heading = 0.5 * np.sin(2 * np.pi * t / 29.0)
    ang = np.zeros((N_ROWS, T))
    vx, vz = speed * np.cos(heading), speed * np.sin(heading)
    ang[PELVIS_TX] = np.cumsum(vx) / sample_freq
    ang[PELVIS_TZ] = np.cumsum(vz) / sample_freq
    ang[PELVIS_TY] = 0.95 + 0.015 * np.cos(2 * ph)
    ang[PELVIS_ROTATION] = heading
    ang[PELVIS_TILT] = 0.03 * np.sin(2 * ph)
    ang[PELVIS_LIST] = 0.02 * np.sin(ph)
    for side, off in (("R", 0.0), ("L", np.pi)):
        hip, add, rot, knee, ank = {
            "R": (HIP_FLEXION_R, HIP_ADDUCTION_R, HIP_ROTATION_R, KNEE_ANG_R, ANKLE_ANG_R),
            "L": (HIP_FLEXION_L, HIP_ADDUCTION_L, HIP_ROTATION_L, KNEE_ANG_L, ANKLE_ANG_L)}[side]
        ang[hip] = 0.15 + 0.4 * np.sin(ph + off)
        ang[add] = -0.05 + 0.04 * np.sin(ph + off + 0.5)
        ang[rot] = 0.03 * np.sin(ph + off)
        ang[knee] = -0.6 + 0.5 * np.cos(ph + off + 0.8)
        ang[ank] = 0.1 * np.sin(ph + off + 1.5)
    ang[LUMBAR_EXTENSION] = -0.03 + 0.01 * np.sin(2 * ph)
    ang[LUMBAR_BENDING] = 0.02 * np.sin(ph)
    ang[LUMBAR_ROTATION] = 0.04 * np.sin(ph + 0.3)
    ang += 1e-4 * rng.standard_normal(ang.shape).cumsum(axis=1) / np.sqrt(np.arange(1, T + 1))
    vel = np.gradient(ang, axis=1) * sample_freq
    return ang, vel


class Loco3dReferenceTrajectories(BaseReferenceTrajectories):
    """Constructor as reference loco3d:22-33; ``path`` / ``control_freq`` are new keyword arguments."""

    def __init__(self, qpos_indices, qvel_indices, adaptations=None, path=None, control_freq=100):
        self._path = path
        super().__init__(500, control_freq, qpos_indices, qvel_indices, adaptations=adaptations)

    def _load_ref_trajecs(self):
        if self._path is None:
            return synthetic_loco3d()
        if self._path.endswith(".npz"):
            z = np.load(self._path)
            return np.asarray(z["angJoi"], np.float64), np.asarray(z["angDJoi"], np.float64)
        import scipy.io as spio
        data = spio.loadmat(self._path, squeeze_me=True)          # loco3d:39-46
        self._qlabels = data["rowNameIK"]
        return np.asarray(data["angJoi"], np.float64), np.asarray(data["angDJoi"], np.float64)

    def _get_COM_Z_pos_index(self):
        return PELVIS_TY

    def tables(self) -> MocapTables:
        T = self._trajec_len
        ref = np.concatenate([gather(self._qpos_full, self._qpos_indices),
                              gather(self._qvel_full, self._qvel_indices)], axis=1)
        # desired velocity = mean of pelvis x / z velocity over the next 0.5 s (loco3d:58-68) -> prefix sums
        window = int(0.5 * self._sample_freq)
        pre = np.zeros((T + 1, 2))
        pre[1:, 0] = np.cumsum(self._qvel_full[PELVIS_TX])
        pre[1:, 1] = np.cumsum(self._qvel_full[PELVIS_TZ])
        return MocapTables(cursor_mode=CURSOR_WRAP, increment=self._increment, ref=ref,
                           step_off=np.zeros(1, np.int32), step_len=np.array([T], np.int32),
                           left_step=np.zeros(1, np.uint8), step_vel=np.zeros(1), step_last_comx=np.zeros(1),
                           com_z_col=self._qpos_indices.index(PELVIS_TY), des_vel_prefix=pre, des_vel_window=window,
                           des_vel_rows=np.stack([self._qvel_full[PELVIS_TX], self._qvel_full[PELVIS_TZ]]))
