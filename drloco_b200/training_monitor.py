"""TrainingMonitor on the batched env (module named training_monitor; reference: common/callback.py) — SURVEY.md §8f "next" row 2 (beyond the hot-path scope; host-side Python only).

Restates reference ``drloco/common/callback.py``: the logging cadence and scalar tag names of ``_on_step`` /
``log_to_tb`` (callback.py:62-170), the model-saving thresholds of ``save_model_if_good`` (:240-266) and the
deterministic evaluation with checkpoint keep / delete of ``eval_walking`` (:272-390).  What differs, and why:

* the Monitor attributes come from the device-side Monitor state through ``env.get_attr`` (one small D2H copy per
  logged step) instead of from N Python envs;
* evaluation runs the ``eval_n_times`` episodes side by side in one small batched env (``ppo.evaluate_walking``)
  instead of one after the other in a reloaded single env; the checkpoint is still written first and then kept
  (renamed with the walked distances) or deleted by the reference's rule;
* scalars go to a writer object with ``add_scalar(tag, value, step)`` / ``add_histogram(tag, values, step, bins)``:
  TensorBoard's SummaryWriter fits; the default writes JSON lines.  W&B is not contacted (no network, not installed);
  the two W&B histograms are written through the same writer under the reference's keys;
* the module-level ``EVAL_INTERVAL`` global the reference mutates is an instance attribute here.
"""
from __future__ import annotations

import json
import os
from typing import Callable, List, Optional

import numpy as np

# callback.py:18-23
EVAL_MORE_FREQUENT_THRES = 3.2e6
EVAL_INTERVAL_RARE = 400e3
EVAL_INTERVAL_FREQUENT = 200e3
EVAL_INTERVAL_MOST_FREQUENT = 100e3


class JsonlWriter:
    """minimal scalar / histogram sink: one JSON object per line."""

    def __init__(self, path: str):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self._f = open(path, "a")

    def add_scalar(self, tag, value, step):
        self._f.write(json.dumps({"tag": tag, "value": float(value), "step": int(step)}) + "\n")

    def add_histogram(self, tag, values, step, bins=40):
        values = np.asarray(values, dtype=np.float64).ravel()
        counts, edges = np.histogram(values, bins=bins) if values.size else (np.zeros(0), np.zeros(0))
        self._f.write(json.dumps({"tag": tag, "step": int(step), "counts": counts.tolist(),
                                  "edges": edges.tolist()}) + "\n")

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()


class TrainingMonitor:
    """Call ``on_training_start()``, then ``on_step()`` after every ``env.step`` of the training loop (what SB3 does with
    a BaseCallback), then ``on_training_end()``.  ``model`` needs ``policy``, ``save(path)`` and ``env``;
    ``env`` needs ``num_envs``, ``get_attr``, ``set_attr``, ``save(path)``."""

    def __init__(self, model, cfg, save_path: str, writer=None, verbose: int = 0,
                 evaluator: Optional[Callable] = None, debug: bool = False):
        self.model, self.cfg, self.save_path, self.verbose, self.debug = model, cfg, save_path, verbose, debug
        self.env = model.env
        self.n_envs = int(self.env.num_envs)
        self.writer = writer
        self._own_writer = writer is None
        self._evaluator = evaluator
        self.num_timesteps = 0
        # callback.py:12-16
        self.max_return = cfg.ep_dur_max * 1 * cfg.rew_scale
        self.ep_return_increment = 0.1 * self.max_return
        self.mean_rew_increment = 0.1 * cfg.rew_scale
        self.eval_interval = EVAL_INTERVAL_RARE if not debug else 10e3
        # callback.py:28-50: save / evaluation bookkeeping and the metrics of the last evaluation
        self.times_surpassed_ep_return_threshold = self.times_surpassed_mean_reward_threshold = 0
        self.n_steps_after_eval = self.eval_interval           # -> the first logging call evaluates
        self.n_saved_models = 0
        self.moved_distances: List[float] = []
        for name in ("mean_walked_distance", "min_walked_distance", "mean_episode_duration", "min_episode_duration",
                     "mean_walking_speed", "min_walking_speed", "mean_reward_means", "count_stable_walks",
                     "summary_score"):
            setattr(self, name, 0)
        self.has_reached_stable_walking = False
        self.steps_to_convergence: Optional[int] = None
        self.failed_eval_runs_indices: List[int] = []
        self.skip_n_steps, self.skipped_steps = 100, 99        # log every 101st call; the second call is the first that logs
        self.saved: List[str] = []          # checkpoints kept on disk (model paths)
        self._last_ep_lens_reset_mio = -1      # the reference also empties the list on its very first call

    # -- callback.py:52-60
    def on_training_start(self) -> None:
        for sub in ("models", "envs", "tb_logs"):
            os.makedirs(os.path.join(self.save_path, sub), exist_ok=True)
        if self.writer is None:
            self.writer = JsonlWriter(os.path.join(self.save_path, "tb_logs", "PPO_1_OWN_LOGS.jsonl"))

    def on_training_end(self) -> None:
        if self.writer is not None and self._own_writer:
            self.writer.close()

    # -- callback.py:62-122
    def on_step(self) -> bool:
        self.num_timesteps += self.n_envs
        # distribution of the episode lengths of the last ~1M steps, not of the whole training.  The reference tests
        # `num_timesteps % 1e6 < 1000` with 8 envs per call (callback.py:69-70); with thousands of envs per call that
        # window would be skipped most of the time, so the crossing of a 1M boundary is tracked explicitly.
        mio = int(self.num_timesteps // 1e6)
        if mio > self._last_ep_lens_reset_mio:
            self._last_ep_lens_reset_mio = mio
            self.env.set_attr("ep_lens", [])
        self.n_steps_after_eval += 1 * self.n_envs
        if self.skipped_steps < self.skip_n_steps:
            self.skipped_steps += 1
            return True
        if self.n_steps_after_eval >= self.eval_interval and not self.debug:
            self.n_steps_after_eval = 0
            walking_stably = self.eval_walking()
            if walking_stably and not self.has_reached_stable_walking:
                self.steps_to_convergence = self.num_timesteps          # wandb summary 'steps_to_convergence'
                self.log_scalar("log_steps_to_convergence", self.num_timesteps)
                self.has_reached_stable_walking = True
            if self.mean_walked_distance >= 20:
                self.eval_interval = EVAL_INTERVAL_RARE
            elif self.mean_walked_distance >= 10:
                self.eval_interval = EVAL_INTERVAL_MOST_FREQUENT
            elif self.mean_walked_distance >= 5:
                self.eval_interval = EVAL_INTERVAL_FREQUENT
        ep_len = self.get_mean("ep_len_smoothed")
        ep_ret = self.get_mean("ep_ret_smoothed")
        mean_rew = self.get_mean("mean_reward_smoothed")
        # no logging during the first episode
        if ep_len < {400: 60, 200: 30, 50: 8, 100: 15}[self.cfg.ctrl_freq]:
            return True
        if not self.debug:
            self.log_to_tb(mean_rew, ep_len, ep_ret)
        if ep_len > 1500:
            self.save_model_if_good(mean_rew, ep_ret)
        self.skipped_steps = 0
        return True

    def get_mean(self, attribute_name):                   # callback.py:125-130
        try:
            return float(np.mean(self.env.get_attr(attribute_name)))
        except Exception:
            return 0.333

    def log_scalar(self, tag, value):                     # callback.py:133-135
        self.writer.add_scalar(tag, value, self.num_timesteps)

    def log_to_tb(self, mean_rew, ep_len, ep_ret):        # callback.py:138-237 (tag names kept verbatim)
        c, n, g = self.cfg, self.n_envs, self.get_mean
        scalars = [
            ("_det_eval/1. Summary Score []", self.summary_score),
            ("_det_eval/2. stable walks count []", self.count_stable_walks),
            ("_det_eval/4. mean eval distance [m]", self.mean_walked_distance),
            ("_det_eval/5. MIN eval distance [m]", self.min_walked_distance),
            ("_det_eval/3. mean step reward [%]", self.mean_reward_means),
            ("_det_eval/6. mean episode duration [%]", self.mean_episode_duration),
            ("_det_eval/7. mean walking speed [m/s]", self.mean_walking_speed),
            ("_train/1. moved distance [m]", g("moved_distance")),
            ("_train/2. episode length [%] (smoothed 0.75)", ep_len / c.ep_dur_max),
            ("_train/3. step reward [] (smoothed 0.25)", (mean_rew - c.alive_bonus) / c.rew_scale),
            ("_train/4. episode return [%] (smoothed 0.75)",
             (ep_ret - ep_len * c.alive_bonus) / (c.ep_dur_max * c.rew_scale)),
        ] + [(f"_rews/{k}. mean ep {name} rew ({n}envs, smoothed 0.9)", g(f"mean_ep_{name}_rew_smoothed"))
             for k, name in ((1, "pos"), (2, "vel"), (3, "com"))]
        for tag, value in scalars:
            self.log_scalar(tag, value)
        lens = [x for per_env in self.env.get_attr("ep_lens") for x in per_env]
        self.writer.add_histogram("_hist/ep_lens", lens, self.num_timesteps, bins=40)
        self.writer.add_histogram("_det_eval/1. walked distances", self.moved_distances, self.num_timesteps, bins=20)

    # -- callback.py:240-266
    def _checkpoint_paths(self, checkpoint: str):
        return (os.path.join(self.save_path, "models", f"model_{checkpoint}.zip"),
                os.path.join(self.save_path, "envs", f"env_{checkpoint}"))

    def _save(self, checkpoint: str):                      # utils.save_model, utils.py:175-192
        model_path, env_path = self._checkpoint_paths(checkpoint)
        self.model.save(model_path)
        self.env.save(env_path)
        return model_path, env_path

    def save_model_if_good(self, mean_rew, ep_ret):
        if self.debug:
            return
        ep_ret_thres = 0.6 * self.max_return + int(self.ep_return_increment *
                                                   (self.times_surpassed_ep_return_threshold + 1))
        if ep_ret > ep_ret_thres:
            path, _ = self._save("ep_ret" + str(ep_ret_thres) + f"_{int(self.num_timesteps / 1e6)}M")
            self.saved.append(path)
            self.times_surpassed_ep_return_threshold += 1
        mean_rew = (mean_rew - self.cfg.alive_bonus) / self.cfg.rew_scale
        mean_rew_thres = 0.4 + self.mean_rew_increment * (self.times_surpassed_mean_reward_threshold + 1)
        if mean_rew > mean_rew_thres:
            self.times_surpassed_mean_reward_threshold += 1     # the reference only counts here (its save is commented out)

    # -- callback.py:272-390
    def eval_walking(self) -> bool:
        cfg = self.cfg
        checkpoint = f"{int(self.num_timesteps / 1e5)}"
        model_path, env_path = self._save(checkpoint)
        eval_n_times = cfg.eval_n_times if self.num_timesteps > 1e6 else 10
        if self._evaluator is not None:
            res = self._evaluator(self.model.policy, self.env, eval_n_times)
        else:
            from .ppo import evaluate_walking
            res = evaluate_walking(self.model.policy, self.env, n_episodes=eval_n_times,
                                   min_stable_distance=cfg.min_stable_distance)
        moved_distances = np.asarray(res["moved_distances"], dtype=np.float64)
        ep_durs = np.asarray(res["ep_durs"], dtype=np.float64)
        mean_rewards = np.asarray(res["mean_rewards"], dtype=np.float64)
        mean_com_x_vels = moved_distances / (ep_durs / cfg.ctrl_freq)
        self.moved_distances = moved_distances.tolist()
        self.mean_walked_distance = float(np.mean(moved_distances))
        self.min_walked_distance = float(np.min(moved_distances))
        self.mean_episode_duration = float(np.mean(ep_durs) / cfg.ep_dur_max)
        self.min_episode_duration = float(np.min(ep_durs))
        self.mean_walking_speed = float(np.mean(mean_com_x_vels))
        self.min_walking_speed = float(np.min(mean_com_x_vels))
        self.mean_reward_means = float((np.mean(mean_rewards) - cfg.alive_bonus) / cfg.rew_scale)
        min_required_distance = cfg.min_stable_distance
        runs_below_min_distance = np.where(moved_distances < min_required_distance)[0]
        count_runs_reached_min_distance = eval_n_times - len(runs_below_min_distance)
        runs_no_falling = np.where((ep_durs == cfg.ep_dur_max) & (moved_distances >= 0.5 * min_required_distance))[0]
        if eval_n_times == cfg.eval_n_times:
            self.failed_eval_runs_indices = runs_below_min_distance.tolist()
        self.count_stable_walks = max(count_runs_reached_min_distance, len(runs_no_falling))
        dt = self.eval_interval / (EVAL_INTERVAL_RARE if self.num_timesteps < EVAL_MORE_FREQUENT_THRES
                                   else EVAL_INTERVAL_FREQUENT)
        self.summary_score += dt * 4 * self.mean_reward_means ** 2 * (self.count_stable_walks / cfg.eval_n_times) ** 4
        were_enough_models_saved = self.n_saved_models >= 5
        walks_humanlike = self.mean_reward_means >= 0.5 * (1 + self.n_saved_models / 10)
        min_dist, mean_dist = int(self.min_walked_distance), int(self.mean_walked_distance)
        is_stable_humanlike_walking = self.count_stable_walks == eval_n_times and walks_humanlike
        retain_model = is_stable_humanlike_walking and not were_enough_models_saved
        if retain_model:
            dists = f"_min{min_dist}mean{mean_dist}"
            new_model_path = model_path[:-4] + dists + ".zip"
            os.rename(model_path, new_model_path)
            os.rename(env_path, env_path + dists)
            self.n_saved_models += 1
            self.saved.append(new_model_path)
        else:
            os.remove(model_path)
            os.remove(env_path)
        if self.verbose:
            print(f"[eval @ {self.num_timesteps}] {'kept' if retain_model else 'deleted'} checkpoint {checkpoint}: "
                  f"min {min_dist} m, mean {mean_dist} m, stable walks {self.count_stable_walks}/{eval_n_times}, "
                  f"mean step reward {self.mean_reward_means:.3f}")
        return bool(is_stable_humanlike_walking)
