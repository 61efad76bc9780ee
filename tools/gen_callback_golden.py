#!/usr/bin/env python
"""Generate tests/golden/callback_trace.json by driving the reference's OWN TrainingMonitor (drloco/common/callback.py,
loaded unmodified from /root/reference) over scripted fake objects.  Runs only in the build container.

What is real: every line of callback.py (cadence, thresholds, tag names, evaluation arithmetic, checkpoint keep / delete /
rename) and utils.save_model.  What is scripted: the training env's Monitor attributes, the evaluation environment
(episodes with a given length, walked distance and step reward, handed out step by step exactly as SB3's VecNormalize
would: normalised rewards, auto-reset) and the policy.  What is stubbed: SB3's BaseCallback / PPO.load, wandb and the
tensorboard SummaryWriter (all three only record what they are given).

tests/test_training_monitor.py replays the same scenarios through drloco_b200.training_monitor.TrainingMonitor.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from baseline import ref_runner as rr                       # noqa: E402

TRAIN_ATTRS = dict(ep_len_smoothed=2000.0, ep_ret_smoothed=2500.0, mean_reward_smoothed=0.9, moved_distance=12.0,
                   mean_ep_pos_rew_smoothed=0.7, mean_ep_vel_rew_smoothed=0.6, mean_ep_com_rew_smoothed=0.5)
EP_LENS = [[100, 200], [300], [], [], [], [], [], []]
RET_VAR = 4.0

# scenario = (name, training-env attribute overrides, num_timesteps at the calls, evaluation episodes of each evaluation,
#             indices of the calls before which n_steps_after_eval is set to the evaluation interval (stands for the
#             50 000 uneventful calls between two evaluations))
# an evaluation episode = (length in control steps incl. the terminal one, walked distance, raw step reward)
SCENARIOS = [
    ("falls_after_2m", {}, [8, 16] + list(range(24, 24 + 101 * 8, 8)), [[(300, 2.0, 0.6)] * 10], []),
    ("stable_walker", {}, [8, 16], [[(3000, 25.3, 0.95)] * 10], []),
    ("interval_7m", {}, [8, 16], [[(500, 7.0, 0.5)] * 10], []),
    ("interval_12m", {}, [8, 16], [[(500, 12.0, 0.5)] * 10], []),
    ("interval_30m", {}, [8, 16], [[(500, 30.0, 0.5)] * 10], []),
    ("short_first_episodes", {"ep_len_smoothed": 12.0}, [8, 16], [[(20, 1.0, 0.3)] * 10], []),
    # beyond 1M steps: 20 evaluation episodes, mixed outcomes (failed-run indices, no-falling rule, summary-score weight)
    ("after_1M_mixed", {}, [2_000_000, 2_000_008],
     [[(3000, 16.0, 0.8)] * 12 + [(3000, 9.0, 0.7)] * 3 + [(700, 4.0, 0.5)] * 2 + [(1500, 15.5, 0.9)] * 3], []),
    # beyond 3.2M steps the score weight changes (EVAL_MORE_FREQUENT_THRES); three evaluations: kept, deleted, kept with
    # the raised human-likeness bar (n_saved_models / 10) and the adapted interval in the score weight
    ("after_3p2M_three_evals", {}, [4_000_000, 4_000_008] + [4_000_008 + 8 * (i + 1) for i in range(202)],
     [[(3000, 22.0, 0.9)] * 20, [(3000, 12.0, 0.85)] * 19 + [(100, 1.0, 0.4)], [(3000, 21.0, 0.78)] * 20], [102, 203]),
]


class Recorder:
    def __init__(self):
        self.events = []

    def scalar(self, tag, value, step):
        self.events.append(dict(kind="scalar", tag=tag, value=float(value), step=int(step)))

    def wandb_log(self, payload, step=None):
        for k, v in payload.items():
            if isinstance(v, dict) and "counts" in v:
                self.events.append(dict(kind="hist", tag=k, counts=v["counts"], edges=v["edges"], step=int(step)))
            else:
                self.events.append(dict(kind="wandb", tag=k, value=float(v), step=None if step is None else int(step)))


class FakeTrainEnv:
    def __init__(self, attrs):
        self.attrs, self.cleared = attrs, 0

    def get_attr(self, name):
        if name == "ep_lens":
            return EP_LENS
        return [self.attrs[name]] * 8

    def set_attr(self, name, value):
        assert name == "ep_lens" and value == []
        self.cleared += 1

    def save(self, path):
        open(path, "w").write("env")


class FakeModel:
    def __init__(self, env):
        self._env = env

    def get_env(self):
        return self._env

    def save(self, path):
        open(path, "w").write("model")


class FakeEvalEnv:
    """one env behind a VecNormalize: normalised rewards, auto-reset, walked distance readable from the inner env"""

    def __init__(self, episodes):
        self.episodes, self.k, self.t = list(episodes), 0, 0
        self.mimic = types.SimpleNamespace(activate_evaluation=lambda: None, get_walked_distance=lambda: self.walked)
        self.venv = types.SimpleNamespace(envs=[types.SimpleNamespace(env=self.mimic)])
        self.ret_rms = types.SimpleNamespace(var=RET_VAR)
        self.walked = 0.0

    def reset(self):
        return np.zeros((1, 29), np.float32)

    def step(self, action):
        n, dist, rew = self.episodes[self.k]
        self.t += 1
        done = self.t >= n
        if done:                       # the env has already been reset when done is reported (callback.py:303-304)
            self.k, self.t, self.walked = self.k + 1, 0, 0.0
        else:
            self.walked = dist * self.t / (n - 1)
        r = np.array([rew / np.sqrt(RET_VAR + 1e-8)])
        return np.zeros((1, 29), np.float32), r, np.array([done]), [{}]


def main():
    rr.load_reference()
    sb3c = rr._stub("stable_baselines3.common.callbacks")

    class BaseCallback:
        def __init__(self, verbose=0):
            self.verbose, self.num_timesteps, self.model, self.training_env = verbose, 0, None, None
    sb3c.BaseCallback = BaseCallback
    from drloco.common import callback as cb
    from drloco.config import hypers as cfg
    rec_holder = {}

    class Writer:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, tag, value, step):
            rec_holder["rec"].scalar(tag, value, step)

        def close(self):
            pass
    cb.SummaryWriter = Writer
    wb = sys.modules["wandb"]
    wb.Histogram = lambda np_histogram: dict(counts=[int(c) for c in np_histogram[0]],
                                             edges=[float(e) for e in np_histogram[1]])
    wb.log = lambda payload, step=None: rec_holder["rec"].wandb_log(payload, step)
    summary = {}
    wb.run = types.SimpleNamespace(summary=summary)
    cb.wandb = wb
    out = dict(meta="reference drloco/common/callback.py TrainingMonitor (unmodified) over scripted env / policy objects; "
                    "n_envs=%d ep_dur_max=%d" % (cfg.n_envs, cfg.ep_dur_max),
               train_attrs=TRAIN_ATTRS, ep_lens=EP_LENS, ret_var=RET_VAR, scenarios=[])
    assert cfg.n_envs == 8
    for name, over, timesteps, evals, force in SCENARIOS:
        save = tempfile.mkdtemp() + "/"
        for sub in ("models", "envs"):
            os.makedirs(save + sub)
        cfg.save_path = save
        cb.EVAL_INTERVAL = cb.EVAL_INTERVAL_RARE              # module global the callback adapts (callback.py:83,98-103)
        summary.clear()
        rec = rec_holder["rec"] = Recorder()
        env = FakeTrainEnv(dict(TRAIN_ATTRS, **over))
        mon = cb.TrainingMonitor()
        mon.model, mon.training_env = FakeModel(env), env
        queue = [list(e) for e in evals]
        eval_results = []

        def load_env(checkpoint, path, env_id, queue=queue):
            return FakeEvalEnv(queue.pop(0))
        cb.utils.load_env = load_env
        cb.PPO = types.SimpleNamespace(load=lambda path: types.SimpleNamespace(
            predict=lambda obs, deterministic: (np.zeros((1, 8), np.float32), None)))
        mon._on_training_start()
        calls = []
        for i, ts in enumerate(timesteps):
            if i in force:
                mon.n_steps_after_eval = cb.EVAL_INTERVAL
            mon.num_timesteps = ts
            n_before = len(rec.events)
            evals_before = len(queue)
            assert mon._on_step() is True
            evaluated = len(queue) < evals_before
            if evaluated:
                eval_results.append(dict(moved_distances=[float(x) for x in mon.moved_distances]))
            calls.append(dict(num_timesteps=ts, n_events=len(rec.events) - n_before, evaluated=evaluated,
                              skipped_steps=mon.skipped_steps, n_steps_after_eval=float(mon.n_steps_after_eval),
                              eval_interval=float(cb.EVAL_INTERVAL), cleared=env.cleared,
                              models=sorted(os.listdir(save + "models")), envs=sorted(os.listdir(save + "envs"))))
        mon._on_training_end()
        state = {k: (float(getattr(mon, k)) if not isinstance(getattr(mon, k), (list, bool)) else getattr(mon, k))
                 for k in ("times_surpassed_ep_return_threshold", "times_surpassed_mean_reward_threshold",
                           "n_saved_models", "mean_walked_distance", "min_walked_distance", "mean_episode_duration",
                           "min_episode_duration", "mean_walking_speed", "min_walking_speed", "mean_reward_means",
                           "count_stable_walks", "summary_score", "has_reached_stable_walking",
                           "failed_eval_runs_indices")}
        assert not queue, "scenario %s: %d scripted evaluations were not run" % (name, len(queue))
        out["scenarios"].append(dict(name=name, attr_overrides=over, timesteps=timesteps, force_eval_calls=force,
                                     eval_episodes=[[list(e) for e in ev] for ev in evals], calls=calls,
                                     events=rec.events, final=state, wandb_summary=dict(summary),
                                     eval_results=eval_results))
        print(name, "events", len(rec.events), "models", calls[-1]["models"], "score", state["summary_score"])
    path = os.path.join(REPO, "tests", "golden", "callback_trace.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
