"""SB3 checkpoint exchange (drloco_b200/checkpoint.py): key mapping of the reference policy, zip layout, tolerant reading
of a pickled VecNormalize.  The SB3-side fixtures are synthesised here from the class / key names SB3 1.0 uses; no SB3
install is available to cross-check (stated in the module docstring)."""
import io
import pickle
import sys
import types
import zipfile

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from drloco_b200 import checkpoint as ck  # noqa: E402
from drloco_b200.ppo import ActorCritic  # noqa: E402


def _sb3_like_state_dict(obs_dim=29, act_dim=8, hidden=(512, 512), shared=True):
    g = torch.Generator().manual_seed(0)
    sd = {"log_std": torch.randn(act_dim, generator=g)}
    d = obs_dim
    for k, h in enumerate(hidden):
        w, b = torch.randn(h, d, generator=g), torch.randn(h, generator=g)
        sd[f"mlp_extractor.policy_net.{2 * k}.weight"], sd[f"mlp_extractor.policy_net.{2 * k}.bias"] = w, b
        sd[f"mlp_extractor.value_net.{2 * k}.weight"] = w if shared else w + 1
        sd[f"mlp_extractor.value_net.{2 * k}.bias"] = b
        d = h
    sd["action_net.weight"], sd["action_net.bias"] = torch.randn(act_dim, d, generator=g), torch.randn(act_dim, generator=g)
    sd["value_net.weight"], sd["value_net.bias"] = torch.randn(1, d, generator=g), torch.randn(1, generator=g)
    return sd


def test_policy_key_mapping_round_trip(tmp_path):
    sd = _sb3_like_state_dict()
    pol = ck.policy_from_sb3_state_dict(ActorCritic(29, 8), sd)
    assert torch.equal(pol.trunk[2].weight, sd["mlp_extractor.policy_net.2.weight"])
    assert torch.equal(pol.action_net.bias, sd["action_net.bias"]) and torch.equal(pol.log_std, sd["log_std"])
    back = ck.policy_to_sb3_state_dict(pol)
    assert sorted(back) == sorted(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    # a checkpoint whose two hidden stacks differ is not the reference's shared-trunk policy
    with pytest.raises(ValueError):
        ck.policy_from_sb3_state_dict(ActorCritic(29, 8), _sb3_like_state_dict(shared=False))
    ck.policy_from_sb3_state_dict(ActorCritic(29, 8), _sb3_like_state_dict(shared=False), strict=False)
    # zip layout of save_to_zip_file: data (json), policy.pth, policy.optimizer.pth, version file
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    path = str(tmp_path / "model_7.zip")
    ck.save_sb3_zip(path, pol, opt, data={"gamma": 0.995})
    with zipfile.ZipFile(path) as z:
        assert sorted(z.namelist()) == ["_stable_baselines3_version", "data", "policy.optimizer.pth", "policy.pth"]
        inner = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=True)
        assert sorted(inner) == sorted(sd)
    pol2 = ActorCritic(29, 8)
    assert ck.load_sb3_zip(path, pol2) == {"gamma": 0.995}
    for (k1, v1), (k2, v2) in zip(pol.state_dict().items(), pol2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)                           # parameters are bit-identical
    x = torch.randn(5, 29, generator=torch.Generator().manual_seed(1))
    # outputs only to rounding: CPU GEMM blocking may depend on the buffers' alignment and the thread count, and with
    # these N(0,1) test weights the 512-term sums are ~20 in magnitude (summation-order differences 1e-5 .. 1e-3
    # absolute); evaluated in float64 the two copies agree to 1e-9
    assert torch.allclose(pol2(x)[0], pol(x)[0], rtol=1e-3, atol=1e-2)
    assert torch.allclose(pol2(x)[1], pol(x)[1], rtol=1e-3, atol=1e-2)
    xd = x.double()
    assert torch.allclose(pol2.double()(xd)[0], pol.double()(xd)[0], rtol=0, atol=1e-9)


def _pickle_sb3_like_vecnormalize(path, D=29):
    """what SB3 1.0 `VecNormalize.save` writes: the wrapper object (class names below) minus venv / class_attributes / ret."""
    mods = {}
    for name in ("stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.vec_env",
                 "stable_baselines3.common.vec_env.vec_normalize", "stable_baselines3.common.running_mean_std",
                 "gym", "gym.spaces", "gym.spaces.box"):
        mods[name] = types.ModuleType(name)
    VN = type("VecNormalize", (), {"__module__": "stable_baselines3.common.vec_env.vec_normalize"})
    RMS = type("RunningMeanStd", (), {"__module__": "stable_baselines3.common.running_mean_std"})
    Box = type("Box", (), {"__module__": "gym.spaces.box"})
    mods["stable_baselines3.common.vec_env.vec_normalize"].VecNormalize = VN
    mods["stable_baselines3.common.running_mean_std"].RunningMeanStd = RMS
    mods["gym.spaces.box"].Box = Box
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        rng = np.random.default_rng(1)
        o, r, box = RMS(), RMS(), Box()
        o.mean, o.var, o.count = rng.standard_normal(D), rng.uniform(0.5, 2, D), 12345.0001
        r.mean, r.var, r.count = np.float64(0.7), np.float64(3.1), 12345.0001
        box.low, box.high, box.shape, box.dtype = -np.ones(D), np.ones(D), (D,), np.dtype("float32")
        vn = VN()
        vn.__dict__.update(obs_rms=o, ret_rms=r, clip_obs=10.0, clip_reward=10.0, gamma=0.99, epsilon=1e-8,
                           training=True, norm_obs=True, norm_reward=True, observation_space=box, action_space=box,
                           num_envs=8, old_obs=np.zeros((8, D)), old_reward=np.zeros(8))
        with open(path, "wb") as f:
            pickle.dump(vn, f)
        return o, r
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_read_and_write_sb3_vecnormalize_pickle(tmp_path):
    path = str(tmp_path / "env_7")
    o, r = _pickle_sb3_like_vecnormalize(path)
    assert "stable_baselines3" not in sys.modules
    with pytest.raises((ImportError, AttributeError)):       # plain pickle cannot load it here ...
        pickle.load(open(path, "rb"))
    sd = ck.read_sb3_vecnormalize(path)                      # ... the tolerant reader can
    np.testing.assert_array_equal(sd["obs_mean"], o.mean)
    np.testing.assert_array_equal(sd["obs_var"], o.var)
    assert sd["obs_count"] == o.count and sd["ret_var"] == 3.1 and sd["ret_mean"] == 0.7 and sd["gamma"] == 0.99
    assert sd["clip_obs"] == 10.0 and sd["norm_reward"] is True and sd["training"] is True
    # writer -> reader round trip, and the stream names SB3's classes
    out = str(tmp_path / "env_out")
    ck.write_sb3_vecnormalize(out, sd, num_envs=4096)
    assert "stable_baselines3" not in sys.modules
    raw = open(out, "rb").read()
    assert b"stable_baselines3.common.vec_env.vec_normalize" in raw and b"RunningMeanStd" in raw
    sd2 = ck.read_sb3_vecnormalize(out)
    for k in sd:
        np.testing.assert_array_equal(np.asarray(sd2[k]), np.asarray(sd[k]), err_msg=k)
