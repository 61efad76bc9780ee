"""Concrete walkers: which mocap rows feed which qpos/qvel entry, index sets, mirror tables.

Mirror of reference drloco/mujoco/mimic_walker3d.py, mimic_walker_165cm_65kg.py and drloco/mujoco/config.py.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

from . import config as cfgm
from .model import WalkerModel, get_model
from .ref_trajecs import straight_walk_trajecs as sw
from .ref_trajecs import loco3d_trajecs as l3
from .ref_trajecs.base_ref_trajecs import MocapTables


def w3d_qpos_indices(n_rows: int = 38) -> List[int]:
    """reference mimic_walker3d.py:11-16."""
    tx, ty, tz = sw.trunk_euler_rows(n_rows)
    return [sw.COM_POSX, sw.COM_POSY, sw.COM_POSZ, tx, ty, tz,
            sw.HIP_SAG_ANG_R, sw.HIP_FRONT_ANG_R, sw.KNEE_ANG_R, sw.ANKLE_ANG_R,
            sw.HIP_SAG_ANG_L, sw.HIP_FRONT_ANG_L, sw.KNEE_ANG_L, sw.ANKLE_ANG_L]


def w3d_qvel_indices() -> List[int]:
    """reference mimic_walker3d.py:18-23."""
    return [sw.COM_VELX, sw.COM_VELY, sw.COM_VELZ, sw.TRUNK_ANGVEL_X, sw.TRUNK_ANGVEL_Y, sw.TRUNK_ANGVEL_Z,
            sw.HIP_SAG_ANGVEL_R, sw.HIP_FRONT_ANGVEL_R, sw.KNEE_ANGVEL_R, sw.ANKLE_ANGVEL_R,
            sw.HIP_SAG_ANGVEL_L, sw.HIP_FRONT_ANGVEL_L, sw.KNEE_ANGVEL_L, sw.ANKLE_ANGVEL_L]


# reference mimic_walker_165cm_65kg.py:6-15
W165_QPOS_INDICES = [l3.PELVIS_TX, l3.PELVIS_TZ, l3.PELVIS_TY, l3.PELVIS_LIST, l3.PELVIS_TILT, l3.PELVIS_ROTATION,
                     l3.LUMBAR_BENDING, l3.LUMBAR_EXTENSION, l3.LUMBAR_ROTATION,
                     l3.HIP_FLEXION_R, l3.HIP_ADDUCTION_R, l3.HIP_ROTATION_R, l3.KNEE_ANG_R, l3.ANKLE_ANG_R,
                     l3.HIP_FLEXION_L, l3.HIP_ADDUCTION_L, l3.HIP_ROTATION_L, l3.KNEE_ANG_L, l3.ANKLE_ANG_L]
W165_QVEL_INDICES = W165_QPOS_INDICES

# mirror tables (reference mimic_env.py:440-489)
MIRROR_OBS_IDX = [0, 1, 2, 3, 4, 5, 6, 11, 12, 13, 14, 7, 8, 9, 10, 15, 16, 17, 18, 19, 20, 25, 26, 27, 28, 21, 22, 23, 24]
MIRROR_OBS_NEG = [2, 4, 6, 8, 12, 16, 18, 20, 22, 26]
MIRROR_ACT_IDX = [4, 5, 6, 7, 0, 1, 2, 3]
MIRROR_ACT_NEG = [1, 5]


@dataclasses.dataclass
class WalkerSpec:
    cfg: cfgm.EnvConfig
    model: WalkerModel
    mocap: MocapTables
    com_indices: List[int]                 # _get_COM_indices
    trunk_rot_indices: List[int]           # _get_trunk_rot_joint_indices
    phase_joints: List[int]                # get_joint_indices_for_phase_estimation
    phase_from_cursor: bool                # mimic_env.py:416-419
    n_des_vel: int
    mirror: bool

    @property
    def obs_dim(self) -> int:
        n_phase = 1 if self.phase_from_cursor else 2 * len(self.phase_joints)
        return n_phase + self.n_des_vel + (self.model.nv - 1) + self.model.nv

    @property
    def act_dim(self) -> int:
        return self.model.nu

    @property
    def frame_skip(self) -> int:
        return self.cfg.frame_skip

    def mirror_tables(self):
        """(obs_idx, obs_sign, act_idx, act_sign); identity when mirroring is off."""
        D, A = self.obs_dim, self.act_dim
        oi, osn = np.arange(D, dtype=np.int32), np.ones(D, np.float32)
        ai, asn = np.arange(A, dtype=np.int32), np.ones(A, np.float32)
        if self.mirror:
            oi = np.array(MIRROR_OBS_IDX, np.int32)
            osn[MIRROR_OBS_NEG] = -1.0
            ai = np.array(MIRROR_ACT_IDX, np.int32)
            asn[MIRROR_ACT_NEG] = -1.0
        return oi, osn, ai, asn


def speed_profile(speeds, speed_profile_duration, control_freq) -> np.ndarray:
    """Desired walking speed per control step, as MimicEnv.activate_speed_control builds it (mimic_env.py:313-322):
    the profile is split into len(speeds)-1 regions of equal length, each a linspace between neighbouring speeds."""
    speeds = list(speeds)
    n_sections = len(speeds) - 1
    assert n_sections >= 1, "a speed profile needs at least two speeds"
    region = int(speed_profile_duration * control_freq / n_sections)
    return np.concatenate([np.linspace(speeds[i], speeds[i + 1], region) for i in range(n_sections)])


def make_spec(cfg: Optional[cfgm.EnvConfig] = None, mocap_path: Optional[str] = None) -> WalkerSpec:
    """env id -> everything the device needs (reference drloco/mujoco/config.py:9-10 ``env_map``)."""
    cfg = cfg or cfgm.EnvConfig()
    model = get_model(cfg.env_id)
    if cfg.env_id == cfgm.STRAIGHT_WALKER:
        rows, _ = sw.load_steps(mocap_path or sw.PATH_CONSTANT_SPEED)
        refs = sw.StraightWalkingTrajectories(w3d_qpos_indices(rows.shape[0]), w3d_qvel_indices(), path=mocap_path)
        return WalkerSpec(cfg, model, refs.tables(), [0, 1, 2], [3, 4, 5], [6, 8, 10, 12], True, 1,
                          cfg.is_mod(cfgm.MOD_MIRR_POLICY))
    if cfg.env_id == cfgm.WALKER_165:
        refs = l3.Loco3dReferenceTrajectories(W165_QPOS_INDICES, W165_QVEL_INDICES, {}, path=mocap_path,
                                              control_freq=cfg.ctrl_freq)
        return WalkerSpec(cfg, model, refs.tables(), [0, 1, 2], [3, 4, 5], [9, 12, 14, 17], False, 2, False)
    raise KeyError(cfg.env_id)
