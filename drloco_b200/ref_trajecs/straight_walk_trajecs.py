"""Straight-walking mocap (400 Hz, split into single steps) -> device tables.

Mirror of reference drloco/ref_trajecs/straight_walk_trajecs.py; row-index constants follow :59-91 for the
constant-speed file (trunk euler rows 35-37, no GRF rows) and for the ramp file (GRF rows 35-36, trunk euler 37-39).
"""
from __future__ import annotations

import os

import numpy as np

from .base_ref_trajecs import (BaseReferenceTrajectories, MocapTables, CURSOR_STEPWISE, DATA_DIR, gather)

PATH_CONSTANT_SPEED = os.path.join(DATA_DIR, "straight_walking_constant_speed_400hz.npz")

# joint position rows (straight:59-63)
COM_POSX, COM_POSY, COM_POSZ = range(0, 3)
TRUNK_ROT_Q1, TRUNK_ROT_Q2, TRUNK_ROT_Q3, TRUNK_ROT_Q4 = range(3, 7)
HIP_FRONT_ANG_R, HIP_SAG_ANG_R, KNEE_ANG_R, ANKLE_ANG_R = range(7, 11)
HIP_FRONT_ANG_L, HIP_SAG_ANG_L, KNEE_ANG_L, ANKLE_ANG_L = range(11, 15)
# joint velocity rows (straight:66-70)
COM_VELX, COM_VELY, COM_VELZ = range(15, 18)
TRUNK_ANGVEL_X, TRUNK_ANGVEL_Y, TRUNK_ANGVEL_Z = range(18, 21)
HIP_FRONT_ANGVEL_R, HIP_SAG_ANGVEL_R, KNEE_ANGVEL_R, ANKLE_ANGVEL_R = range(21, 25)
HIP_FRONT_ANGVEL_L, HIP_SAG_ANGVEL_L, KNEE_ANGVEL_L, ANKLE_ANGVEL_L = range(25, 29)
FOOT_POSX_L, FOOT_POSY_L, FOOT_POSZ_L, FOOT_POSX_R, FOOT_POSY_R, FOOT_POSZ_R = range(29, 35)


def trunk_euler_rows(n_rows: int):
    """(TRUNK_ROT_X, Y, Z): rows 35-37 in the 38-row constant-speed file, 37-39 in the 40-row ramp file (straight:85-91)."""
    return (35, 36, 37) if n_rows == 38 else (37, 38, 39)


TRUNK_ROT_X, TRUNK_ROT_Y, TRUNK_ROT_Z = trunk_euler_rows(38)


def smooth_exponential(data, alpha=0.9):
    """reference drloco/common/utils.py:264-268."""
    smoothed = np.array(data, dtype=np.float64, copy=True)
    for t in range(1, len(data)):
        smoothed[t] = alpha * data[t] + (1 - alpha) * smoothed[t - 1]
    return smoothed


def _seq_mean(x) -> float:
    acc = 0.0
    for v in x.tolist():
        acc += v
    return acc / len(x)


def load_steps(path: str):
    """-> (rows float64 [R, T], step_len int32 [n_steps]) from a compiled .npz or the reference's .mat."""
    if path.endswith(".npz"):
        z = np.load(path)
        return np.asarray(z["rows"], np.float64), np.asarray(z["step_len"], np.int32)
    import scipy.io as spio
    data = spio.loadmat(path, squeeze_me=True)["Data"].flatten()
    steps = [np.asarray(s, dtype=np.float64) for s in data]
    return np.concatenate(steps, axis=1), np.array([s.shape[1] for s in steps], np.int32)


class StraightWalkingTrajectories(BaseReferenceTrajectories):
    """Constructor as reference straight:98-104 (the reference drops ``adaptations`` too, Q11); ``path`` is new."""

    def __init__(self, qpos_indices, q_vel_indices, adaptations=None, mirror_refs=False, path=None):
        if mirror_refs:
            raise NotImplementedError("mirror_refs is never enabled in the reference (SURVEY.md Q10)")
        self._path = path or PATH_CONSTANT_SPEED
        super().__init__(400, 200, qpos_indices, q_vel_indices)

    def _load_ref_trajecs(self):
        self.rows, self.step_len = load_steps(self._path)
        # qvel is read from the same matrix as qpos (Q7, straight:155,165-166,320)
        return self.rows, self.rows

    def _get_COM_Z_pos_index(self):
        return COM_POSZ

    def tables(self) -> MocapTables:
        rows, step_len = self.rows, self.step_len
        off = np.concatenate([[0], np.cumsum(step_len)[:-1]]).astype(np.int32)
        n = len(step_len)
        seg = [rows[:, off[i]:off[i] + step_len[i]] for i in range(n)]
        # left step <=> swing knee is the left one (straight:221-230)
        left = np.array([np.max(s[KNEE_ANGVEL_L]) > np.max(s[KNEE_ANGVEL_R]) for s in seg], np.uint8)
        # per-step walking speed, exponentially smoothed (straight:393-415)
        # (the reference averages Python floats of an object array: plain left-to-right summation)
        speeds = smooth_exponential([_seq_mean(s[COM_VELX, :]) for s in seg], alpha=0.2)
        last_x = np.array([s[COM_POSX, -1] for s in seg], np.float64)
        first_x = np.array([s[COM_POSX, 0] for s in seg])
        if not np.all(first_x < 0.005):          # straight:343-345
            raise AssertionError("The COM X Position on each new step trajectory should start with 0.0")
        ref = np.concatenate([gather(rows, self._qpos_indices), gather(rows, self._qvel_indices)], axis=1)
        return MocapTables(cursor_mode=CURSOR_STEPWISE, increment=self._increment, ref=ref, step_off=off,
                           step_len=step_len.astype(np.int32), left_step=left, step_vel=speeds,
                           step_last_comx=last_x, com_z_col=self._qpos_indices.index(COM_POSZ))


def synthetic_straight_rows(n_steps=30, seed=0, sample_freq=400.0):
    """A smooth periodic gait with the constant-speed file's 38-row layout, for boxes without the mocap."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(249, 280, size=n_steps)
    segs = []
    for i, L in enumerate(lens):
        t = np.arange(L) / L
        s = np.zeros((38, L))
        sw, st = (1, 0) if i % 2 else (0, 1)       # odd steps swing left
        speed = 1.45 + 0.02 * rng.standard_normal()
        s[COM_POSX] = speed * np.arange(L) / sample_freq
        s[COM_POSY] = 0.01 * np.sin(np.pi * t) * (1 if i % 2 else -1)
        s[COM_POSZ] = 1.03 + 0.012 * np.cos(2 * np.pi * t)
        s[COM_VELX] = speed
        s[COM_VELY] = np.gradient(s[COM_POSY]) * sample_freq
        s[COM_VELZ] = np.gradient(s[COM_POSZ]) * sample_freq
        hip = (HIP_SAG_ANG_R, HIP_SAG_ANG_L)
        knee = (KNEE_ANG_R, KNEE_ANG_L)
        ank = (ANKLE_ANG_R, ANKLE_ANG_L)
        s[hip[sw]] = -0.35 + 0.7 * t
        s[hip[st]] = 0.35 - 0.7 * t
        s[knee[sw]] = 0.15 + 0.9 * np.sin(np.pi * t) ** 2
        s[knee[st]] = 0.12 + 0.1 * np.sin(np.pi * t)
        s[ank[sw]] = 0.05 * np.sin(2 * np.pi * t)
        s[ank[st]] = 0.1 - 0.25 * t
        for a, v in ((HIP_SAG_ANG_R, HIP_SAG_ANGVEL_R), (HIP_SAG_ANG_L, HIP_SAG_ANGVEL_L),
                     (KNEE_ANG_R, KNEE_ANGVEL_R), (KNEE_ANG_L, KNEE_ANGVEL_L),
                     (ANKLE_ANG_R, ANKLE_ANGVEL_R), (ANKLE_ANG_L, ANKLE_ANGVEL_L)):
            s[v] = np.gradient(s[a]) * sample_freq
        s[TRUNK_ROT_Q1] = 1.0
        s[36] = 0.1       # trunk euler y (sagittal lean), rows 35-37
        segs.append(s)
    return np.concatenate(segs, axis=1), lens.astype(np.int32)
