"""Checkpoint exchange with the reference's Stable-Baselines3 files — SURVEY.md §8f "next" row 3 (host-side, no GPU).

The reference saves two files per checkpoint (``drloco/common/utils.py:175-184``): ``models/model_<ckpt>.zip``
(``PPO.save``: a zip with ``data`` (json), ``policy.pth``, ``policy.optimizer.pth``, ``pytorch_variables.pth``,
``_stable_baselines3_version``) and ``envs/env_<ckpt>`` (``VecNormalize.save``: ``pickle.dump`` of the wrapper without
its venv).  Neither stable-baselines3 nor gym is installed in the build image, so

* READING works without them: the zip's ``policy.pth`` is a plain torch state dict whose keys follow the reference's
  ``CustomActorCriticPolicy`` (``drloco/custom/policies.py:13-80``: ``mlp_extractor.policy_net`` and ``.value_net`` are
  the SAME layers — Q15 — so both names hold the same tensors); the VecNormalize pickle is read with an unpickler that
  substitutes attribute bags for the SB3 / gym classes and only numpy has to be importable;
* WRITING produces files of the same structure: the zip can be given to ``model.set_parameters(path)`` of an SB3 model
  built by the reference code, the env file to ``VecNormalize.load(path, venv)``.  Both writers were checked against the
  readers above and against the SB3 1.0 sources from memory only - NOT against a real SB3 install (none available).
"""
from __future__ import annotations

import io
import json
import pickle
import sys
import types
import zipfile
from typing import Dict, Optional

import numpy as np
import torch

SB3_VERSION = "1.0"
_VN_MODULE = "stable_baselines3.common.vec_env.vec_normalize"
_RMS_MODULE = "stable_baselines3.common.running_mean_std"


# ---------------------------------------------------------------------------------------------------------------------
# policy  <->  SB3 ``policy.pth``
# ---------------------------------------------------------------------------------------------------------------------
def _trunk_linear_indices(policy) -> list:
    return [i for i, m in enumerate(policy.trunk) if isinstance(m, torch.nn.Linear)]


def policy_to_sb3_state_dict(policy) -> Dict[str, torch.Tensor]:
    """``ActorCritic`` -> state dict with the key names of the reference's SB3 policy."""
    sd = {"log_std": policy.log_std.detach().clone()}
    for i in _trunk_linear_indices(policy):
        for part in ("weight", "bias"):
            t = getattr(policy.trunk[i], part).detach().clone()
            sd[f"mlp_extractor.policy_net.{i}.{part}"] = t
            sd[f"mlp_extractor.value_net.{i}.{part}"] = t          # shared layers: same tensor under both names
    for name in ("action_net", "value_net"):
        for part in ("weight", "bias"):
            sd[f"{name}.{part}"] = getattr(getattr(policy, name), part).detach().clone()
    return sd


def policy_from_sb3_state_dict(policy, sd: Dict[str, torch.Tensor], strict: bool = True):
    """load an SB3 ``policy.pth`` state dict (reference policy layout) into an ``ActorCritic``."""
    own = {"log_std": sd["log_std"]}
    for i in _trunk_linear_indices(policy):
        for part in ("weight", "bias"):
            p, v = sd[f"mlp_extractor.policy_net.{i}.{part}"], sd.get(f"mlp_extractor.value_net.{i}.{part}")
            if strict and v is not None and not torch.equal(p, v):
                raise ValueError("policy and value hidden layers differ: not a shared-trunk checkpoint of the reference "
                                 "policy (policies.py:38-41); pass strict=False to take the policy branch")
            own[f"trunk.{i}.{part}"] = p
    for name in ("action_net", "value_net"):
        for part in ("weight", "bias"):
            own[f"{name}.{part}"] = sd[f"{name}.{part}"]
    policy.load_state_dict(own)
    return policy


def save_sb3_zip(path: str, policy, optimizer: Optional[torch.optim.Optimizer] = None, data: Optional[dict] = None):
    """write ``model_<ckpt>.zip`` in the layout of SB3's ``save_to_zip_file``."""
    def blob(obj):
        b = io.BytesIO()
        torch.save(obj, b)
        return b.getvalue()
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps(data or {}))
        z.writestr("policy.pth", blob(policy_to_sb3_state_dict(policy)))
        # SB3's set_parameters(exact_match=True) expects the optimiser entry: always present (empty state when none given)
        z.writestr("policy.optimizer.pth", blob(optimizer.state_dict() if optimizer is not None
                                                else {"state": {}, "param_groups": []}))
        z.writestr("_stable_baselines3_version", SB3_VERSION)


def load_sb3_zip(path: str, policy, strict: bool = True, map_location="cpu") -> dict:
    """read ``policy.pth`` of an SB3 model zip into ``policy``; returns the zip's json ``data`` (hyper-parameters)."""
    with zipfile.ZipFile(path) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location=map_location, weights_only=True)
        data = json.loads(z.read("data").decode()) if "data" in z.namelist() else {}
    policy_from_sb3_state_dict(policy, sd, strict)
    return data


# ---------------------------------------------------------------------------------------------------------------------
# VecNormalize statistics  <->  SB3 ``VecNormalize.save`` pickle
# ---------------------------------------------------------------------------------------------------------------------
class _Bag:
    """stands in for any class the pickle names but this image cannot import (SB3, gym)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:                                    # (dict, slots) form
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)


# what a pickled VecNormalize legitimately needs besides the SB3 / gym classes (which are replaced by _Bag): numpy's
# array reconstruction helpers and a handful of plain containers.  Nothing callable with side effects (eval, exec,
# os.system, ...) can be reached through find_class.
_SAFE_GLOBALS = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy", "float64"), ("numpy", "float32"), ("numpy", "int64"),
    ("numpy", "bool_"), ("numpy.core.numeric", "_frombuffer"), ("numpy._core.numeric", "_frombuffer"),
    ("collections", "OrderedDict"), ("collections", "deque"), ("copyreg", "_reconstructor"),
    ("_codecs", "encode"),
    ("builtins", "object"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"),
    ("builtins", "frozenset"), ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"),
    ("builtins", "bytes"), ("builtins", "bytearray"), ("builtins", "complex"), ("builtins", "slice"),
}


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _SAFE_GLOBALS:
            return super().find_class(module, name)
        root = module.split(".")[0]
        if root in ("builtins", "os", "posix", "nt", "subprocess", "sys", "importlib", "shutil", "socket"):
            raise pickle.UnpicklingError(f"refusing to load {module}.{name} from an env file")
        return type(name, (_Bag,), {"__module__": module})


def read_sb3_vecnormalize(path: str) -> dict:
    """statistics of a pickled SB3 ``VecNormalize`` as a ``B200VecNormalize.load_state_dict`` payload."""
    with open(path, "rb") as f:
        vn = _TolerantUnpickler(f).load()
    g = vn.__dict__
    D = int(np.asarray(g["obs_rms"].mean).size)
    return {"obs_mean": np.asarray(g["obs_rms"].mean, np.float64).reshape(D),
            "obs_var": np.asarray(g["obs_rms"].var, np.float64).reshape(D),
            "obs_count": float(g["obs_rms"].count),
            "ret_mean": float(np.asarray(g["ret_rms"].mean)), "ret_var": float(np.asarray(g["ret_rms"].var)),
            "ret_count": float(g["ret_rms"].count),
            "clip_obs": float(g.get("clip_obs", 10.0)), "clip_reward": float(g.get("clip_reward", 10.0)),
            "gamma": float(g.get("gamma", 0.99)), "epsilon": float(g.get("epsilon", 1e-8)),
            "norm_obs": bool(g.get("norm_obs", True)), "norm_reward": bool(g.get("norm_reward", True)),
            "training": bool(g.get("training", True))}


def write_sb3_vecnormalize(path: str, sd: dict, num_envs: int = 1) -> None:
    """pickle the statistics under SB3's class names (what ``VecNormalize.__getstate__`` leaves: no venv, no
    class_attributes, no ret).  Spaces are left out: ``VecNormalize.load`` -> ``set_venv`` takes them from the venv."""
    created = []

    def stub(module, name):
        parts = module.split(".")
        for k in range(1, len(parts) + 1):
            m = ".".join(parts[:k])
            if m not in sys.modules:
                sys.modules[m] = types.ModuleType(m)
                created.append(m)
        cls = type(name, (), {"__module__": module})
        setattr(sys.modules[module], name, cls)
        return cls

    had = {m: getattr(sys.modules.get(m), n, None) for m, n in ((_VN_MODULE, "VecNormalize"), (_RMS_MODULE, "RunningMeanStd"))}
    try:
        VN, RMS = stub(_VN_MODULE, "VecNormalize"), stub(_RMS_MODULE, "RunningMeanStd")

        def rms(mean, var, count):
            r = RMS()
            r.mean, r.var, r.count = mean, var, float(count)
            return r
        vn = VN()
        vn.__dict__.update(
            obs_rms=rms(np.asarray(sd["obs_mean"], np.float64), np.asarray(sd["obs_var"], np.float64), sd["obs_count"]),
            ret_rms=rms(np.float64(sd["ret_mean"]), np.float64(sd["ret_var"]), sd["ret_count"]),
            clip_obs=float(sd.get("clip_obs", 10.0)), clip_reward=float(sd.get("clip_reward", 10.0)),
            gamma=float(sd.get("gamma", 0.99)), epsilon=float(sd.get("epsilon", 1e-8)),
            training=bool(sd.get("training", True)), norm_obs=bool(sd.get("norm_obs", True)),
            norm_reward=bool(sd.get("norm_reward", True)), num_envs=int(num_envs),
            old_obs=np.array([]), old_reward=np.array([]))
        with open(path, "wb") as f:
            pickle.dump(vn, f)
    finally:
        for (m, n), old in zip(((_VN_MODULE, "VecNormalize"), (_RMS_MODULE, "RunningMeanStd")), had.values()):
            if m in sys.modules and m not in created:
                if old is None:
                    delattr(sys.modules[m], n)
                else:
                    setattr(sys.modules[m], n, old)
        for m in created:
            sys.modules.pop(m, None)
