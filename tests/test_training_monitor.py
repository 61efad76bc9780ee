"""TrainingMonitor host logic (reference drloco/common/callback.py) on fake model / env objects: cadence, tag names,
save thresholds, evaluation bookkeeping and checkpoint keep / delete.  No GPU."""
import json
import os

import numpy as np
import pytest

from drloco_b200.config import EnvConfig
from drloco_b200.training_monitor import (EVAL_INTERVAL_FREQUENT, EVAL_INTERVAL_MOST_FREQUENT, EVAL_INTERVAL_RARE,
                                          JsonlWriter, TrainingMonitor)


class _Env:
    num_envs = 8

    def __init__(self):
        self.attrs = dict(ep_len_smoothed=2000.0, ep_ret_smoothed=2500.0, mean_reward_smoothed=0.9, moved_distance=12.0,
                          mean_ep_pos_rew_smoothed=0.7, mean_ep_vel_rew_smoothed=0.6, mean_ep_com_rew_smoothed=0.5)
        self.cleared = 0

    def get_attr(self, name):
        if name == "ep_lens":
            return [[100, 200], [300]] + [[] for _ in range(self.num_envs - 2)]
        return [self.attrs[name]] * self.num_envs

    def set_attr(self, name, value):
        assert name == "ep_lens" and value == []
        self.cleared += 1

    def save(self, path):
        open(path, "w").write("env")


class _Model:
    policy = object()

    def __init__(self):
        self.env = _Env()

    def save(self, path):
        open(path, "w").write("model")


def _evaluator(dist, dur, rew):
    def run(policy, env, n):
        return dict(moved_distances=[dist] * n, ep_durs=[dur] * n, mean_rewards=[rew] * n)
    return run


def _records(path):
    return [json.loads(line) for line in open(path)]


def test_cadence_tags_and_checkpoint_rules(tmp_path):
    cfg = EnvConfig()
    sp = str(tmp_path) + "/"
    # 1) a walker that falls after 2 m: evaluation on the first call, checkpoint deleted
    mon = TrainingMonitor(_Model(), cfg, sp, evaluator=_evaluator(2.0, 300, 0.6))
    mon.on_training_start()
    assert mon.on_step() is True                                       # skipped_steps 99 -> 100: the first call only counts
    assert mon.num_timesteps == 8 and mon.env.cleared == 1            # 8 % 1e6 < 1000 -> ep_lens emptied
    assert mon.mean_walked_distance == 0 and not os.listdir(sp + "models")
    mon.on_step()                                                      # the second call evaluates and logs
    assert mon.mean_walked_distance == 2.0 and mon.count_stable_walks == 0 and mon.summary_score == 0
    assert abs(mon.mean_reward_means - (0.6 - cfg.alive_bonus) / cfg.rew_scale) < 1e-12
    assert abs(mon.mean_walking_speed - 2.0 / (300 / cfg.ctrl_freq)) < 1e-12
    assert mon.failed_eval_runs_indices == []                          # only recorded for full 20-episode evaluations
    assert os.listdir(sp + "models") == ["model_ep_ret2100.0_0M.zip"]  # eval checkpoint removed; ep_ret 2500 > 0.6*3000+300
    assert os.listdir(sp + "envs") == ["env_ep_ret2100.0_0M"]
    assert mon.times_surpassed_ep_return_threshold == 1 and mon.times_surpassed_mean_reward_threshold == 1
    assert mon.eval_interval == EVAL_INTERVAL_RARE
    # the next 100 calls are skipped, the 101st logs again
    mon.writer.flush()
    n0 = len(_records(sp + "tb_logs/PPO_1_OWN_LOGS.jsonl"))
    for _ in range(100):
        mon.on_step()
    mon.writer.flush()
    assert len(_records(sp + "tb_logs/PPO_1_OWN_LOGS.jsonl")) == n0
    mon.on_step()
    mon.on_training_end()
    recs = _records(sp + "tb_logs/PPO_1_OWN_LOGS.jsonl")
    assert len(recs) == 2 * n0
    tags = [r["tag"] for r in recs[:n0]]
    assert tags == ["_det_eval/1. Summary Score []", "_det_eval/2. stable walks count []",
                    "_det_eval/4. mean eval distance [m]", "_det_eval/5. MIN eval distance [m]",
                    "_det_eval/3. mean step reward [%]", "_det_eval/6. mean episode duration [%]",
                    "_det_eval/7. mean walking speed [m/s]", "_train/1. moved distance [m]",
                    "_train/2. episode length [%] (smoothed 0.75)", "_train/3. step reward [] (smoothed 0.25)",
                    "_train/4. episode return [%] (smoothed 0.75)", "_rews/1. mean ep pos rew (8envs, smoothed 0.9)",
                    "_rews/2. mean ep vel rew (8envs, smoothed 0.9)", "_rews/3. mean ep com rew (8envs, smoothed 0.9)",
                    "_hist/ep_lens", "_det_eval/1. walked distances"]
    by = {r["tag"]: r for r in recs[:n0]}
    assert abs(by["_train/2. episode length [%] (smoothed 0.75)"]["value"] - 2000 / 3000) < 1e-12
    assert abs(by["_train/3. step reward [] (smoothed 0.25)"]["value"] - 0.7) < 1e-12
    assert abs(by["_train/4. episode return [%] (smoothed 0.75)"]["value"] - (2500 - 2000 * 0.2) / 3000) < 1e-12
    assert sum(by["_hist/ep_lens"]["counts"]) == 3 and by["_hist/ep_lens"]["step"] == 16

    # 2) a stable, human-like walker: checkpoint kept and renamed, training may stop, evaluation interval stays rare
    sp2 = str(tmp_path) + "/good/"
    mon = TrainingMonitor(_Model(), cfg, sp2, evaluator=_evaluator(25.3, 3000, 0.95))
    mon.on_training_start()
    mon.on_step(), mon.on_step()
    assert mon.has_reached_stable_walking and mon.steps_to_convergence == 16
    assert mon.count_stable_walks == 10                                  # 10 episodes before 1M steps
    assert "model_0_min25mean25.zip" in os.listdir(sp2 + "models") and "env_0_min25mean25" in os.listdir(sp2 + "envs")
    assert mon.n_saved_models == 1
    assert abs(mon.summary_score - 1.0 * 4 * 0.75 ** 2 * (10 / 20) ** 4) < 1e-12
    mon.on_training_end()

    # 3) interval adaptation by mean walked distance (callback.py:98-103)
    for dist, want in ((7.0, EVAL_INTERVAL_FREQUENT), (12.0, EVAL_INTERVAL_MOST_FREQUENT), (30.0, EVAL_INTERVAL_RARE)):
        mon = TrainingMonitor(_Model(), cfg, str(tmp_path) + f"/d{dist}/", evaluator=_evaluator(dist, 500, 0.5))
        mon.on_training_start()
        mon.on_step(), mon.on_step()
        assert mon.eval_interval == want
        mon.on_training_end()

    # 4) short first episodes: nothing is logged, nothing saved (callback.py:110-111)
    mon = TrainingMonitor(_Model(), cfg, str(tmp_path) + "/short/", evaluator=_evaluator(1.0, 20, 0.3))
    mon.env.attrs["ep_len_smoothed"] = 12.0
    mon.on_training_start()
    mon.on_step(), mon.on_step()
    mon.on_training_end()
    assert _records(str(tmp_path) + "/short/tb_logs/PPO_1_OWN_LOGS.jsonl") == []
    assert mon.skipped_steps == 100                                      # not reset -> tries again on the next call


def test_jsonl_writer_histogram(tmp_path):
    w = JsonlWriter(str(tmp_path / "x" / "log.jsonl"))
    w.add_scalar("a", np.float32(1.5), 7)
    w.add_histogram("h", [1, 2, 2, 3], 7, bins=3)
    w.add_histogram("empty", [], 7)
    w.close()
    r = _records(str(tmp_path / "x" / "log.jsonl"))
    assert r[0] == {"tag": "a", "value": 1.5, "step": 7} and r[1]["counts"] == [1, 2, 1] and r[2]["counts"] == []


class _RecordingWriter:
    """collects what the monitor logs in the shape tools/gen_callback_golden.py records the reference's tensorboard /
    wandb calls"""

    def __init__(self):
        self.events = []

    def add_scalar(self, tag, value, step):
        self.events.append(dict(kind="scalar", tag=tag, value=float(value), step=int(step)))

    def add_histogram(self, tag, values, step, bins=40):
        counts, edges = np.histogram(values, bins=bins)
        self.events.append(dict(kind="hist", tag=tag, counts=[int(c) for c in counts], edges=[float(e) for e in edges],
                                step=int(step)))

    def flush(self):
        pass

    def close(self):
        pass


def test_trace_matches_the_reference_callback(tmp_path):
    """tests/golden/callback_trace.json was produced by the reference's own TrainingMonitor (drloco/common/callback.py,
    unmodified) over scripted env / policy objects (tools/gen_callback_golden.py).  The same scenarios through this
    package's TrainingMonitor must log the same tags and values at the same calls, keep / delete / rename the same
    checkpoint files and end in the same bookkeeping state."""
    golden = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "callback_trace.json")))
    cfg = EnvConfig()
    assert cfg.ep_dur_max == 3000 and cfg.ctrl_freq == 200
    ret_scale = np.sqrt(golden["ret_var"] + 1e-8)
    for sc in golden["scenarios"]:
        sp = str(tmp_path / sc["name"]) + "/"
        queue = [list(e) for e in sc["eval_episodes"]]
        results = list(sc["eval_results"])

        def evaluator(policy, env, n, queue=queue, results=results):
            eps = queue.pop(0)
            assert n == len(eps)                                        # 10 evaluation episodes up to 1M steps, then 20
            # what the reference's evaluation loop (callback.py:292-311) derives from the scripted episodes: duration
            # incl. the terminal step, mean of the un-normalised rewards of the non-terminal steps, distance read
            # before the terminal step
            return dict(moved_distances=results.pop(0)["moved_distances"], ep_durs=[e[0] for e in eps],
                        mean_rewards=[float(np.mean([e[2] / ret_scale * ret_scale] * (e[0] - 1))) for e in eps])
        model = _Model()
        model.env.attrs.update(golden["train_attrs"])
        model.env.attrs.update(sc["attr_overrides"])
        writer = _RecordingWriter()
        mon = TrainingMonitor(model, cfg, sp, writer=writer, evaluator=evaluator)
        mon.on_training_start()
        for i, (ts, call) in enumerate(zip(sc["timesteps"], sc["calls"])):
            if i in sc["force_eval_calls"]:
                mon.n_steps_after_eval = mon.eval_interval
            mon.num_timesteps = ts - model.env.num_envs                  # SB3 sets num_timesteps; on_step() adds n_envs
            n_before, evals_before = len(writer.events), len(queue)
            assert mon.on_step() is True
            where = f"{sc['name']} call {i}"
            assert mon.num_timesteps == ts
            assert len(writer.events) - n_before == call["n_events"], where
            assert (len(queue) < evals_before) == call["evaluated"], where
            assert mon.skipped_steps == call["skipped_steps"] and mon.eval_interval == call["eval_interval"], where
            assert mon.n_steps_after_eval == call["n_steps_after_eval"], where
            assert sorted(os.listdir(sp + "models")) == call["models"], where
            assert sorted(os.listdir(sp + "envs")) == call["envs"], where
            # ep_lens: the reference empties the list on every call inside `num_timesteps % 1e6 < 1000`; here once per
            # crossed 1M boundary (deliberate, see on_step) - both have done so by the time the reference has
            assert (model.env.cleared > 0) == (call["cleared"] > 0), where
        mon.on_training_end()
        assert not queue
        assert len(writer.events) == len(sc["events"])
        for got, want in zip(writer.events, sc["events"]):
            assert got["tag"] == want["tag"], sc["name"]
            if want["kind"] == "hist":
                assert got["kind"] == "hist" and got["counts"] == want["counts"] and got["step"] == want["step"]
                np.testing.assert_allclose(got["edges"], want["edges"], rtol=1e-12, atol=1e-12)
            else:                                    # tensorboard scalar, or the one value the reference sends to wandb
                assert got["kind"] == "scalar" and (want["step"] is None or got["step"] == want["step"])
                assert got["value"] == pytest.approx(want["value"], rel=1e-12, abs=1e-15), (sc["name"], got["tag"])
        for name, want in sc["final"].items():
            got = getattr(mon, name)
            if isinstance(want, (list, bool)):
                assert got == want, (sc["name"], name)
            else:
                assert float(got) == pytest.approx(want, rel=1e-12, abs=1e-15), (sc["name"], name)
        if "steps_to_convergence" in sc["wandb_summary"]:
            assert mon.steps_to_convergence == sc["wandb_summary"]["steps_to_convergence"]
        else:
            assert mon.steps_to_convergence is None
