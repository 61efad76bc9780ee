"""baseline/ref_runner.py — runs the reference's OWN environment code (SURVEY.md section 8c recipe).  TEST / BASELINE
INFRASTRUCTURE: used by tools/gen_golden.py (golden vectors) and by ``bench.py --impl reference`` (the CPU arm).

The reference's MimicWalker3dEnv / MimicEnv / Monitor / StraightWalkingTrajectories classes are imported *unmodified*
from a reference tree (``/root/reference`` in the build container, or its copy ``baseline/_ref`` made by
baseline/build_ref.py, which travels to the GPU box).  The third-party packages they import but that are not installed
(gym, mujoco_py, seaborn, matplotlib, wandb, stable_baselines3) are replaced by import stubs, and gym's ``MujocoEnv`` by
a minimal stand-in whose ``sim`` is the float64 physics restatement (oracle/walker_physics.c -> oracle/liboracle.so):
MuJoCo itself is a third-party binary that cannot be installed here.  Two textual substitutions are applied while loading
reference modules, both forced by the checkout rather than chosen: ``PATH_REF_TRAJECS = PATH_CONSTANT_SPEED`` (the
default ramp mocap is missing, Q8) and ``from collections import Iterable`` -> ``collections.abc`` (Python >= 3.10).
"""
import collections
import collections.abc
import os
import sys
import tempfile
import time
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
REF_CONTAINER = "/root/reference"
REF_COPY = os.path.join(REPO, "baseline", "_ref")
REF = REF_CONTAINER          # set by use_reference_tree()

from drloco_b200.model import get_model            # noqa: E402
from oracle.physics import OraclePhysics           # noqa: E402


def reference_tree():
    """the reference tree to import from: the container's checkout if present, else the copy under baseline/_ref"""
    for root in (REF_CONTAINER, REF_COPY):
        if os.path.isdir(os.path.join(root, "drloco", "mujoco")):
            return root
    return None


def use_reference_tree(root):
    global REF
    REF = root


PHYSICS_SECONDS = [0.0]      # time spent inside the physics library (for the physics / Python-glue split)


class _Dummy:
    """absorbs any attribute access / call / item assignment (plot configuration of the reference)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()

    def __setitem__(self, k, v):
        pass

    def __getitem__(self, k):
        return _Dummy()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    def _ga(n):                                # PEP 562
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()
    mod.__getattr__ = _ga
    sys.modules[name] = mod
    return mod


class MujocoException(Exception):
    pass


class _Box:
    def __init__(self, low, high):
        self.low, self.high = np.asarray(low, np.float32), np.asarray(high, np.float32)
        self.shape = self.low.shape

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(np.float32)


class _SimData:
    pass


class _Sim:
    """Stand-in for mujoco_py.MjSim backed by the physics oracle."""

    def __init__(self, model):
        self.phys = OraclePhysics(model)
        self.data = _SimData()
        self.data.qpos = model.qpos0.copy()
        self.data.qvel = np.zeros(model.nv)
        self.data.ctrl = np.zeros(model.nu)
        self.data.actuator_force = np.zeros(model.nu)
        self.data.site_xpos = self.phys.site_xpos(self.data.qpos)
        self.data.time = 0.0
        self._m = model

    def reset(self):
        self.data.qpos[:] = self._m.qpos0
        self.data.qvel[:] = 0
        self.phys.qacc_warm[:] = 0
        self.forward()

    def forward(self):
        self.data.site_xpos = self.phys.site_xpos(self.data.qpos)

    def step(self):
        m = self._m
        self.data.actuator_force[:] = np.clip(np.clip(self.data.ctrl, m.act_ctrlrange[:, 0], m.act_ctrlrange[:, 1])
                                              * m.act_gear, m.act_forcerange[:, 0], m.act_forcerange[:, 1])
        t0 = time.perf_counter()
        bad = self.phys.step(self.data.qpos, self.data.qvel, self.data.ctrl, 1)
        PHYSICS_SECONDS[0] += time.perf_counter() - t0
        if bad:
            raise MujocoException("unstable simulation")


class _ModelView:
    def __init__(self, model):
        self.actuator_ctrlrange = model.act_ctrlrange.copy()
        self.actuator_forcerange = model.act_forcerange.copy()


class FakeMujocoEnv:
    """What gym 0.18.0's MujocoEnv does for MimicEnv, minus rendering."""

    def __init__(self, model_path, frame_skip):
        name = os.path.basename(model_path)
        self._wm = get_model({"walker3d_flat_feet.xml": "StraightMimicWalker",
                              "walker_165cm_65kg.xml": "MimicWalker165cm65kg"}[name])
        self.frame_skip = frame_skip
        self.sim = _Sim(self._wm)
        self.data = self.sim.data
        self.model = _ModelView(self._wm)
        self.init_qpos, self.init_qvel = self.data.qpos.copy(), self.data.qvel.copy()
        cr = self.model.actuator_ctrlrange
        self.action_space = _Box(cr[:, 0], cr[:, 1])
        observation, _reward, done, _info = self.step(self.action_space.sample())
        assert not done
        self.observation_space = _Box(np.full(observation.shape, -np.inf), np.full(observation.shape, np.inf))

    def seed(self, seed=None):
        return [seed]

    def reset(self):
        self.sim.reset()
        return self.reset_model()

    def set_state(self, qpos, qvel):
        self.data.qpos[:] = qpos
        self.data.qvel[:] = qvel
        self.sim.phys.qacc_warm[:] = 0
        self.sim.forward()

    def do_simulation(self, ctrl, n_frames):
        self.data.ctrl[:] = ctrl
        for _ in range(n_frames):
            self.sim.step()

    @property
    def dt(self):
        return self._wm.timestep * self.frame_skip


class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)


def install_stubs():
    collections.Iterable = collections.abc.Iterable
    gym = _stub("gym", Wrapper=_Wrapper)
    gym.utils = _stub("gym.utils", EzPickle=type("EzPickle", (), {"__init__": lambda self, *a, **k: None}))
    _stub("gym.envs")
    _stub("gym.envs.mujoco")
    _stub("gym.envs.mujoco.mujoco_env", MujocoEnv=FakeMujocoEnv)
    mj = _stub("mujoco_py", MjSimState=_Dummy)
    mj.builder = _stub("mujoco_py.builder", MujocoException=MujocoException)
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot", "wandb", "stable_baselines3",
                 "stable_baselines3.common", "stable_baselines3.common.vec_env"):
        _stub(name)
    for cls in ("DummyVecEnv", "SubprocVecEnv", "VecNormalize"):
        setattr(sys.modules["stable_baselines3.common.vec_env"], cls, type(cls, (), {}))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load_reference(hypers_subst=(), mimic_env_subst=()):
    """import the reference package with the two forced substitutions; returns (env class, Monitor class, utils).
    ``hypers_subst``: textual edits of drloco/config/hypers.py (settings the reference expects its user to edit).
    ``mimic_env_subst``: textual edits of drloco/mujoco/mimic_env.py (only tools/gen_golden.py's speed-control fixture
    uses it, for the one token without which that path raises; every other run loads the file as it is)."""
    sys.path.insert(0, REF)
    # is_remote() <=> 'code/torch' in cwd: no viewer, n_envs = 8 (Q12/Q13)
    work = os.path.join(tempfile.mkdtemp(), "code", "torch")
    os.makedirs(work)
    os.chdir(work)
    import torch  # noqa: F401  (hypers.py imports it; load the real one before the stubs go in)
    install_stubs()
    if hypers_subst:
        import drloco.config  # noqa: F401
        _load_module_with("drloco.config.hypers", "drloco/config/hypers.py", list(hypers_subst))
    name = "drloco.ref_trajecs.straight_walk_trajecs"
    path = os.path.join(REF, "drloco/ref_trajecs/straight_walk_trajecs.py")
    src = open(path).read().replace("PATH_REF_TRAJECS = PATH_SPEED_RAMP", "PATH_REF_TRAJECS = PATH_CONSTANT_SPEED")
    import drloco.ref_trajecs  # noqa: F401  (package)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    setattr(sys.modules["drloco.ref_trajecs"], "straight_walk_trajecs", mod)
    if mimic_env_subst:
        import drloco.mujoco  # noqa: F401
        _load_module_with("drloco.mujoco.mimic_env", "drloco/mujoco/mimic_env.py", list(mimic_env_subst))
    from drloco.mujoco.mimic_walker3d import MimicWalker3dEnv
    from drloco.mujoco.monitor_wrapper import Monitor
    from drloco.common import utils
    return MimicWalker3dEnv, Monitor, utils


def _load_module_with(name, rel_path, replacements):
    """exec a reference module from source with textual substitutions and register it (also on its package)."""
    path = os.path.join(REF, rel_path)
    src = open(path).read()
    for a, b in replacements:
        assert a in src, (rel_path, a)
        src = src.replace(a, b)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    pkg, _, leaf = name.rpartition(".")
    setattr(sys.modules[pkg], leaf, mod)
    return mod


