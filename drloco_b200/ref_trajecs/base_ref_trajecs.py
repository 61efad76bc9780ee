"""Base class of the reference-trajectory *table compilers*.

In the reference, ``BaseReferenceTrajectories`` (drloco/ref_trajecs/base_ref_trajecs.py:6-158) is a per-env CPU
cursor over a (rows x samples) matrix: ``next()`` advances ``_pos``, ``get_qpos()/get_qvel()`` gather the rows the
walker needs.  On the B200 the cursor lives in device memory, one per environment, and is advanced inside the fused
step kernel; the classes here keep the reference's names and constructor signatures but only *compile* the mocap
into the flat tables the kernels index:

    ref  float [n_samples, 2*nv]   qpos rows then qvel rows, already gathered in model order (base:44-56)
    step_off / step_len            where each mocap step starts and how long it is (1 step for loco3d)

The per-env cursor logic itself is restated once for tests in ``oracle/env_oracle.py`` and once for the GPU in
``csrc/mimic_step.cu``.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, Optional, Sequence

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")

# cursor advance rules
CURSOR_STEPWISE = 0   # StraightWalkingTrajectories.next (straight_walk_trajecs.py:141-159)
CURSOR_WRAP = 1       # BaseReferenceTrajectories.next (base_ref_trajecs.py:95-103)


@dataclasses.dataclass
class MocapTables:
    """Everything the device (and the oracle) needs to know about one mocap."""
    cursor_mode: int
    increment: int                 # samples per control step (base:87-93)
    ref: np.ndarray                # float64 [n_samples, 2*nv]
    step_off: np.ndarray           # int32 [n_steps]
    step_len: np.ndarray           # int32 [n_steps]
    left_step: np.ndarray          # uint8 [n_steps]  (straight:221-230)
    step_vel: np.ndarray           # float64 [n_steps] smoothed mean COM-X velocity per step (straight:393-415)
    step_last_comx: np.ndarray     # float64 [n_steps] COM-X at the last sample of each step (straight:339)
    com_z_col: int                 # column of ``ref`` holding the COM-Z position (straight:482, loco3d:48)
    des_vel_prefix: Optional[np.ndarray] = None   # float64 [n_samples+1, 2] prefix sums for loco3d:51-68
    des_vel_window: int = 0
    des_vel_rows: Optional[np.ndarray] = None     # float64 [2, n_samples] the velocity rows themselves (the oracle
                                                  # takes np.mean of the slice exactly like loco3d:62-68)

    @property
    def n_steps(self) -> int:
        return int(self.step_len.shape[0])

    @property
    def n_samples(self) -> int:
        return int(self.ref.shape[0])


class BaseReferenceTrajectories:
    """Same constructor as the reference (base_ref_trajecs.py:20-42); subclasses provide ``_load_ref_trajecs``."""

    def __init__(self, sample_freq, control_freq, qpos_indices, qvel_indices, data_labels=(), adaptations=None):
        self._sample_freq = sample_freq
        self._control_freq = control_freq
        self._qpos_indices = list(qpos_indices)
        self._qvel_indices = list(qvel_indices)
        self._qlabels = data_labels
        self._qpos_full, self._qvel_full = self._load_ref_trajecs()
        self._n_joints, self._trajec_len = self._qpos_full.shape
        self.adapt_trajectories(adaptations or {})
        if not (len(data_labels) == 0 or len(data_labels) in (self._n_joints, len(self._qpos_indices))):
            raise AssertionError(
                "Please provide a label for each row in the data matrix.\n"
                f"You provided {len(data_labels)} labels for a matrix of shape {self._qpos_full.shape}.\n")
        self._set_increment()

    def _set_increment(self):
        increment = self._sample_freq / self._control_freq
        if not float(increment).is_integer():
            raise AssertionError(
                "Please check your control frequency and the sample frequency of the reference data!"
                f"The sampling frequency ({self._sample_freq}) of the reference data should be equal to "
                f"or an integer multiple of the control frequency which is set to {self._control_freq}.")
        self._increment = int(increment)

    def adapt_trajectories(self, adaptations_dict: Dict[int, float]):
        """Scale individual rows (base_ref_trajecs.py:105-118)."""
        for index, scalar in adaptations_dict.items():
            self._qpos_full[index, :] *= scalar
            if self._qvel_full is not self._qpos_full:
                self._qvel_full[index, :] *= scalar

    def get_kinematics_labels(self):
        return self._qlabels

    # ---- to override -------------------------------------------------------------------------
    def _load_ref_trajecs(self):
        raise NotImplementedError

    def _get_COM_Z_pos_index(self) -> int:
        raise NotImplementedError

    def tables(self) -> MocapTables:
        raise NotImplementedError


def gather(rows: np.ndarray, idx: Sequence[int]) -> np.ndarray:
    """rows [R, T] -> [T, len(idx)] (the reference's fancy index ``full[indices, pos]`` for every pos)."""
    return np.ascontiguousarray(rows[list(idx), :].T)
