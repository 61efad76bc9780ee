#!/usr/bin/env python
"""baseline/build_ref.py — materialise baseline/_ref, the reference arm of bench.py (`--impl reference`).

rgalljamov/DRLoco is a set of plain Python scripts (no setup.py / pyproject.toml: `pip install --target baseline/_ref
/root/reference` fails with "Directory is not installable"), so the "install" is a verbatim copy of the files the
env-step path needs: the `drloco` package and the one mocap recording present in the checkout.  baseline/_ref is
git-ignored (reference sources never enter the history) but travels to the GPU box with the snapshot, where
/root/reference does not exist.  Nothing is modified; baseline/ref_runner.py imports the copy under import stubs for the
third-party packages that are not installed (gym, mujoco_py, SB3, ...) and drives it over oracle/liboracle.so.

Run in the build container:  python baseline/build_ref.py
"""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")
KEEP = ["drloco", "mocaps/straight_walking/Trajecs_Constant_Speed_400Hz.mat"]


def build(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "drloco")):
        if verbose:
            print(f"{SRC} is not present: baseline/_ref left as it is")
        return os.path.isdir(os.path.join(DST, "drloco"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for rel in KEEP:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    for root, dirs, files in os.walk(DST):          # the reference checkout is read-only; the copy need not be
        for name in dirs + files:
            os.chmod(os.path.join(root, name), 0o755 if name in dirs else 0o644)
    with open(os.path.join(DST, "ORIGIN.txt"), "w") as f:
        f.write("verbatim copy of /root/reference/{drloco, mocaps/straight_walking/Trajecs_Constant_Speed_400Hz.mat}\n"
                "made by baseline/build_ref.py; not part of the repository history\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"baseline/_ref: {n} files copied from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
