#!/usr/bin/env python
"""Compile a DRLoco straight-walking mocap .mat file into the flat .npz table format used by drloco_b200.

Input  (reference format, drloco/ref_trajecs/straight_walk_trajecs.py:304-320): MATLAB v5 file whose ``Data``
        field is an object array of n_steps matrices of shape (n_rows, len_i), 400 Hz.
Output: ``rows``  float64 [n_rows, total_samples]  (all steps concatenated along time),
        ``step_len`` int32 [n_steps], ``sample_freq`` float.

Usage: python tools/compile_mocap.py /root/reference/mocaps/straight_walking/Trajecs_Constant_Speed_400Hz.mat \
           drloco_b200/data/straight_walking_constant_speed_400hz.npz
"""
import sys

import numpy as np
import scipy.io as spio


def main(src: str, dst: str) -> None:
    data = spio.loadmat(src, squeeze_me=True)["Data"].flatten()
    steps = [np.asarray(s, dtype=np.float64) for s in data]
    n_rows = steps[0].shape[0]
    assert all(s.shape[0] == n_rows for s in steps)
    rows = np.concatenate(steps, axis=1)
    step_len = np.array([s.shape[1] for s in steps], np.int32)
    np.savez_compressed(dst, rows=rows, step_len=step_len, sample_freq=np.float64(400.0))
    print(f"{src}: {len(steps)} steps, {n_rows} rows, {rows.shape[1]} samples -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
