"""oracle/env_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU float64 restatement of the reference's environment logic around the physics:

  RefCursor        drloco/ref_trajecs/straight_walk_trajecs.py:141-171,322-348,417-422,460-474 and
                   drloco/ref_trajecs/base_ref_trajecs.py:44-103, loco3d_trajecs.py:51-97
  OracleMimicEnv   drloco/mujoco/mimic_env.py:60-126 (step), :131-139, :142-168, :170-192, :403-437, :440-489, :526-572,
                   :592-649 plus the MujocoEnv calls it makes (do_simulation / set_state / reset, gym 0.18.0)
  OracleMonitor    drloco/mujoco/monitor_wrapper.py:88-166 with drloco/common/utils.py:312-329 (per-process smoothing)
  OracleVecEnv     SB3 1.0 DummyVecEnv.step_wait semantics (auto-reset, terminal_observation)
  OracleVecNormalize / RunningMeanStd   SB3 1.0 VecNormalize as used at drloco/common/utils.py:130-132

Pinned by tests/golden/*.npz, which tools/gen_golden.py produced by running the reference's *unmodified* MimicEnv /
Monitor / ref_trajecs Python under import stubs; tests/test_oracle_golden.py compares bit-for-bit.
Deliberate deviations from the reference (SURVEY.md quirk register): Q4 (in-place mutation of the mocap by
adjust_COM_Z_pos accumulating across episodes) is replaced by a per-episode z offset; RSI draws can be injected.
"""
from __future__ import annotations

import math
import random
from typing import List, Optional

import numpy as np

from drloco_b200.ref_trajecs.base_ref_trajecs import CURSOR_STEPWISE, MocapTables


class RefCursor:
    """Per-env position on the reference motion."""

    def __init__(self, mocap: MocapTables, nv: int):
        self.t = mocap
        self.nv = nv
        self.i_step = 0
        self.pos = 0
        self.len = int(mocap.step_len[0])
        self.count_steps_same_vel = 1      # straight:124 — never reset (Q3)
        self.dist = 0.0                    # offset added to the COM-X column of the current step (Q2)
        self.z_off = 0.0                   # adjust_COM_Z_pos offset of this episode, applies to the RSI step only (Q4 waiver)
        self.rsi_step = 0                  # ``self._step`` of the reference: the step chosen at the last RSI
        self.n_deterministic_inits = 0
        # which step's table `_qpos_full` points at, and which table got the in-place COM-Z shift: the logical step
        # except after a deterministic init, where the reference keeps reading step 0 until the first transition (Q27)
        self.data_step = 0
        self.shift_step = 0

    # -- lookups --------------------------------------------------------------------------------
    def _row(self):
        return self.t.ref[int(self.t.step_off[self.data_step]) + self.pos]

    def get_qpos(self):
        q = self._row()[:self.nv].copy()
        if self.t.cursor_mode == CURSOR_STEPWISE:
            q[0] += self.dist                               # straight:346
            if self.data_step == self.shift_step:
                q[self.t.com_z_col] -= self.z_off           # base:126-127 through the alias of straight:467-469;
                                                            # every other step is an unshifted copy (straight:342)
        else:
            q[self.t.com_z_col] -= self.z_off               # base:126-127 shifts the whole recording
        return q

    def get_qvel(self):
        return self._row()[self.nv:2 * self.nv].copy()

    def is_step_left(self) -> bool:
        return bool(self.t.left_step[self.i_step])          # straight:233-234

    def get_phase_variable(self) -> float:
        return self.pos / self.len                          # straight:173-177

    def get_desired_walking_velocity_vector(self) -> List[float]:
        if self.t.cursor_mode == CURSOR_STEPWISE:           # straight:417-422, 479-480
            return [float(self.t.step_vel[max(0, self.i_step - self.count_steps_same_vel + 1)])]
        end = min(self.pos + self.t.des_vel_window, self.len - 1)   # loco3d:58-68
        n = end - self.pos
        if n <= 0:
            return [math.nan, math.nan]                     # mean of an empty slice (Q23)
        rows = self.t.des_vel_rows                          # np.mean of the slice, as the reference computes it
        return [np.mean(rows[0, self.pos:end]), np.mean(rows[1, self.pos:end])]

    # -- motion ---------------------------------------------------------------------------------
    def next(self):
        inc = self.t.increment
        if self.t.cursor_mode == CURSOR_STEPWISE:           # straight:141-159
            self.pos += inc
            dif = self.pos - self.len + 1
            if dif > 0:
                n = self.t.n_steps                          # straight:322-348
                if self.i_step >= n - 1:
                    self.i_step = 0 if self.t.left_step[self.i_step] else 1
                else:
                    self.i_step += 1
                    self.count_steps_same_vel += 1
                self.dist = float(self.t.step_last_comx[self.rsi_step])   # ``_step`` is never updated (Q2)
                self.len = int(self.t.step_len[self.i_step])
                self.data_step = self.i_step
                self.pos = dif
        else:                                               # base:95-103
            self.pos += inc
            if self.pos >= self.len - 1:
                self.pos = 0

    def init_random(self, i_step: Optional[int] = None, pos: Optional[int] = None):
        """straight:460-474 / base:79-85; (i_step, pos) may be injected."""
        if self.t.cursor_mode == CURSOR_STEPWISE:
            if i_step is None:
                i_step = random.randint(0, self.t.n_steps - 1)
            self.i_step = int(i_step)
            self.len = int(self.t.step_len[self.i_step])
            if pos is None:
                pos = random.randint(0, self.len - 1)
            self.pos = int(pos)
        else:
            self.i_step, self.len = 0, int(self.t.step_len[0])
            self.pos = int(np.random.randint(0, self.len)) if pos is None else int(pos)
        self.data_step = self.shift_step = self.i_step
        self._begin_episode()

    def init_deterministic(self, eval_n_times: int):
        """straight:237-265 / base:69-77."""
        if self.t.cursor_mode == CURSOR_STEPWISE:
            # straight:240 `self.reset()` points `_qpos_full` / `_trajec_len` at step 0 (:163-167) and the lines after it
            # only move `_i_step`, `_step` and `_pos`: data and length stay step 0's until the first transition (Q27)
            self.i_step = self.n_deterministic_inits
            self.len = int(self.t.step_len[0])
            self.pos = int(0.75 * int(self.t.step_len[self.i_step]))
            self.data_step = self.shift_step = 0
            self.n_deterministic_inits += 1
            if self.n_deterministic_inits >= eval_n_times:
                self.n_deterministic_inits = 0
        else:
            self.i_step, self.len, self.pos = 0, int(self.t.step_len[0]), 0
        self._begin_episode()

    def _begin_episode(self):
        self.dist = 0.0
        self.z_off = 0.0
        self.rsi_step = self.i_step

    def adjust_COM_Z_pos(self, offset: float):
        self.z_off += offset


class OracleMimicEnv:
    """One environment: reference MimicEnv + the walker subclass + MujocoEnv glue, over a pluggable physics object
    exposing ``step(q, v, ctrl, nsub) -> blew_up``, ``site_xpos(q)`` (oracle.physics.OraclePhysics)."""

    def __init__(self, spec, physics):
        self.spec, self.cfg, self.phys = spec, spec.cfg, physics
        m = spec.model
        self.nv, self.nu = m.nv, m.nu
        self.refs = RefCursor(spec.mocap, m.nv)
        self.qpos, self.qvel = m.qpos0.copy(), np.zeros(m.nv)
        # gym's Box(actuator_ctrlrange) is float32; with the float32 actions SB3 passes, the scaling below is
        # therefore float32 arithmetic in the reference, which the GPU path reproduces bit for bit
        self.low, self.high = m.act_ctrlrange[:, 0].astype(np.float32), m.act_ctrlrange[:, 1].astype(np.float32)
        self.act_force = np.zeros(m.nu)
        self.pos_rew = self.vel_rew = self.com_rew = 0.0
        self.ep_dur = 0
        self.walked_distance = 0.0
        self._EVAL_MODEL = False
        self._FOLLOW_DESIRED_SPEED_PROFILE = False           # mimic_env.py:33
        self.desired_walking_speed_trajectory = None
        self._PLAYBACK_REF_TRAJECS = False                   # mimic_env.py:266
        self.mirr_obs_idx, self.mirr_obs_sign, self.mirr_act_idx, self.mirr_act_sign = spec.mirror_tables()
        self.last_ctrl = np.zeros(m.nu)

    # mimic_env.py:170-192
    def _rescale_actions(self, action):
        action = np.clip(np.asarray(action), -1, 1)          # keeps the caller's dtype, like the reference
        return np.array([a * self.high[i] if a > 0 else np.abs(a) * self.low[i] for i, a in enumerate(action)])

    def mirror_action(self, acts):                           # mimic_env.py:483-489
        return acts[self.mirr_act_idx] * self.mirr_act_sign

    def mirror_obs(self, obs):                               # mimic_env.py:440-480
        return obs[self.mirr_obs_idx] * self.mirr_obs_sign

    def _excl_com(self, x):
        keep = [i for i in range(self.nv) if i not in self.spec.com_indices]
        return x[keep]

    def get_pose_reward(self):                               # mimic_env.py:592-601
        dif = self._excl_com(self.qpos) - self._excl_com(self.refs.get_qpos())
        return float(np.exp(-3 * np.sum(np.square(dif))))

    def get_vel_reward(self):                                # mimic_env.py:603-611
        dif = self._excl_com(self.qvel) - self._excl_com(self.refs.get_qvel())
        return float(np.exp(-0.05 * np.sum(np.square(dif))))

    def get_com_reward(self):                                # mimic_env.py:613-622
        ci = self.spec.com_indices
        dif = self.qpos[ci] - self.refs.get_qpos()[ci]
        return float(np.exp(-16 * np.sum(np.square(dif))))

    def get_imitation_reward(self):                          # mimic_env.py:633-649
        w_pos, w_vel, w_com, _ = self.cfg.rew_weights
        self.pos_rew, self.vel_rew, self.com_rew = self.get_pose_reward(), self.get_vel_reward(), self.get_com_reward()
        return (w_pos * self.pos_rew + w_vel * self.vel_rew + w_com * self.com_rew) * self.cfg.rew_scale

    def _get_ET_reward(self):                                # mimic_env.py:149-168 with ep_rews == [] always (Q1)
        mean_epret_smoothed = 0.0
        if self.ep_dur >= self.cfg.ep_dur_max:
            mean_step_rew = mean_epret_smoothed / self.ep_dur
            return float(np.sum(mean_step_rew * np.power(self.cfg.gamma, np.arange(self.ep_dur))))
        return -1 * mean_epret_smoothed

    def estimate_phase_vars(self):                           # mimic_env.py:330-401
        out = []
        for j in self.spec.phase_joints:
            pos, vel = self.qpos[j], self.qvel[j]
            out += [np.arctan2(vel, -pos) / np.pi, np.linalg.norm([pos, vel]) / 5]
        return out

    def activate_speed_control(self, speeds=(1.0, 1.0), speed_profile_duration=10):   # mimic_env.py:298-322
        self._FOLLOW_DESIRED_SPEED_PROFILE = True
        n_sections = len(speeds) - 1
        region = int(speed_profile_duration * self.cfg.ctrl_freq / n_sections)
        self.desired_walking_speed_trajectory = np.concatenate(
            [np.linspace(speeds[i], speeds[i + 1], region) for i in range(n_sections)])

    def _get_obs(self):                                      # mimic_env.py:403-437
        if self._FOLLOW_DESIRED_SPEED_PROFILE:               # mimic_env.py:406-408
            # the reference then does `*self.desired_walking_speed` on this scalar and raises TypeError (Q26); the
            # restatement follows the intent: the scalar fills the first desired-velocity slot, the others stay 0
            prof = self.desired_walking_speed_trajectory
            n_des = len(self.refs.get_desired_walking_velocity_vector())
            des = [prof[self.ep_dur % len(prof)]] + [0.0] * (n_des - 1)
        else:
            des = self.refs.get_desired_walking_velocity_vector()
        phases = [self.refs.get_phase_variable()] if self.spec.phase_from_cursor else self.estimate_phase_vars()
        obs = np.array([*phases, *des, *self.qpos[1:], *self.qvel])
        if self.spec.mirror and self.refs.is_step_left():
            obs = self.mirror_obs(obs)
        return obs

    def step(self, action):                                  # mimic_env.py:60-126
        ctrl = self._rescale_actions(action)
        if self.spec.mirror and self.refs.is_step_left():
            ctrl = self.mirror_action(ctrl)
        ctrl = ctrl.astype(np.float64)                       # sim.data.ctrl[:] = action
        self.last_ctrl = ctrl
        m = self.spec.model
        self.act_force = np.clip(np.clip(ctrl, m.act_ctrlrange[:, 0], m.act_ctrlrange[:, 1]) * m.act_gear,
                                 m.act_forcerange[:, 0], m.act_forcerange[:, 1])
        if not self._PLAYBACK_REF_TRAJECS and self.phys.step(self.qpos, self.qvel, ctrl, self.spec.frame_skip):
            obs = self.reset()                               # MujocoException path, mimic_env.py:86-91
            return obs, 0, True, {"blowup": True}
        self.refs.next()
        if self._PLAYBACK_REF_TRAJECS:                       # body of the playback loop, mimic_env.py:273-275,284-293
            self.qpos[:], self.qvel[:] = self.refs.get_qpos(), self.refs.get_qvel()
        obs = self._get_obs()
        self.ep_dur += 1
        vel_vec = np.clip(self.qvel[:2], -5.5, 5.5)          # mimic_env.py:131-139
        self.walked_distance += float(np.linalg.norm(vel_vec)) * 1 / self.cfg.ctrl_freq
        com_z = self.qpos[self.spec.com_indices[-1]]
        done = bool(com_z < self.cfg.fall_z or self.ep_dur >= self.cfg.ep_dur_max)
        self.et_flags = self.do_terminate_early()            # mimic_env.py:122-123 (commented out in the reference)
        if self.cfg.early_termination and self.et_flags[0]:
            done = True
        reward = self._get_ET_reward() if done else self.get_imitation_reward() + self.cfg.alive_bonus
        return obs, reward, done, {}

    def do_terminate_early(self):                            # mimic_env.py:652-702 (3D branch: trunk dofs 3,4,5)
        if self._PLAYBACK_REF_TRAJECS:
            return [False] * 4
        qpos, ref = self.qpos, self.refs.get_qpos()
        com_height, com_y = qpos[2], qpos[1]
        front, sag, _ = qpos[3:6]
        front_dev = abs(qpos[3] - ref[3])
        trunk = bool((sag > 0.3 or sag < -0.05) or front_dev > 0.2)
        drunk = bool(abs(com_y) > 0.2)
        low = bool(com_height < 0.75)
        return [low or trunk or drunk, low, trunk, drunk]

    def reset(self, i_step=None, pos=None):                  # MujocoEnv.reset + mimic_env.py:526-572
        self.ep_dur = 0
        self.walked_distance = 0
        if self._EVAL_MODEL or self._FOLLOW_DESIRED_SPEED_PROFILE:   # mimic_env.py:536-537
            self.refs.init_deterministic(self.cfg.eval_n_times)
        else:
            self.refs.init_random(i_step, pos)
        qpos, qvel = self.refs.get_qpos(), self.refs.get_qvel()
        lowest = float(np.min(self.phys.site_xpos(qpos)[:, 2]))
        qpos[self.spec.com_indices[-1]] -= lowest
        self.refs.adjust_COM_Z_pos(lowest)
        self.qpos[:], self.qvel[:] = qpos, qvel
        if hasattr(self.phys, "qacc_warm"):
            self.phys.qacc_warm[:] = 0
        rew = self.get_imitation_reward()
        assert rew > 0.95 * self.cfg.rew_scale, f"Reward should be around 1 after RSI, but was {rew}!"
        self.refs.next()
        return self._get_obs()

    def get_actuator_torques(self, abs_mean=False):          # mimic_env.py:251-253
        return float(np.mean(np.abs(self.act_force))) if abs_mean else self.act_force.copy()


class OracleMonitor:
    """reference Monitor.step statistics; smoothing state is per env (= per process under SubprocVecEnv, Q17)."""

    def __init__(self, env: OracleMimicEnv):
        self.env = env
        self._ewa = {}
        self.ep_len = 0
        self.rewards: List[float] = []
        self.ep_pos_rews, self.ep_vel_rews, self.ep_com_rews = [], [], []
        self.ep_torques_abs: List[float] = []
        self.ep_lens: List[int] = []
        self.returns: List[float] = []
        self.rsi_positions, self.et_positions, self.difficult_rsi_phases = [], [], []
        self.init_pos = 0
        self.ep_difficult: List[bool] = []                    # per finished episode: did it enter difficult_rsi_phases
        self.ep_len_smoothed = self.ep_ret_smoothed = self.mean_reward_smoothed = 0
        self.mean_ep_pos_rew_smoothed = self.mean_ep_vel_rew_smoothed = self.mean_ep_com_rew_smoothed = 0
        self.mean_abs_ep_torque_smoothed = 0
        self.median_abs_torque_smoothed = 0
        self.moved_distance = 0

    def _smooth(self, label, new_value, smoothing_factor=0.9):   # utils.py:312-329
        if label not in self._ewa:
            self._ewa[label] = new_value
            return new_value
        new_average = smoothing_factor * new_value + (1 - smoothing_factor) * self._ewa[label]
        self._ewa[label] = new_average
        return new_average

    def step(self, action):                                   # monitor_wrapper.py:88-166
        obs, reward, done, info = self.env.step(action)
        if self.ep_len == 0:
            self.init_pos = self.env.refs.pos
            self.rsi_positions.append(self.init_pos)
        self.ep_len += 1
        self.rewards.append(reward)
        self.ep_pos_rews.append(self.env.pos_rew)
        self.ep_vel_rews.append(self.env.vel_rew)
        self.ep_com_rews.append(self.env.com_rew)
        self.ep_torques_abs.append(self.env.get_actuator_torques(True))
        if done:
            self.et_positions.append(self.env.refs.pos)
            ep_rewards = self.rewards[-self.ep_len:]
            if self.ep_len > 1:
                self.mean_reward_smoothed = self._smooth("rew", float(np.mean(ep_rewards[:-1])))
            self.mean_ep_pos_rew_smoothed = self._smooth("ep_pos_rew", float(np.mean(self.ep_pos_rews)))
            self.mean_ep_vel_rew_smoothed = self._smooth("ep_vel_rew", float(np.mean(self.ep_vel_rews)))
            self.mean_ep_com_rew_smoothed = self._smooth("ep_com_rew", float(np.mean(self.ep_com_rews)))
            ep_return = float(np.sum(ep_rewards))
            self.returns.append(ep_return)
            self.ep_ret_smoothed = self._smooth("ep_ret", ep_return, 0.25)
            self.ep_lens.append(self.ep_len)
            self.ep_len_smoothed = self._smooth("ep_len", self.ep_len, 0.75)
            self.ep_difficult.append(bool(self.ep_len < self.ep_len_smoothed * 0.75))
            if self.ep_difficult[-1]:
                self.difficult_rsi_phases.append(self.init_pos)
            self.ep_len = 0
            self.moved_distance = self.env.walked_distance
            self.mean_abs_ep_torque_smoothed = self._smooth("mean_ep_tor", float(np.mean(self.ep_torques_abs)), 0.75)
            self.median_abs_torque_smoothed = self._smooth("med_ep_tor", float(np.median(self.ep_torques_abs)), 0.75)
            self.ep_torques_abs = []
        return obs, reward, done, info


class OracleVecEnv:
    """N monitored envs stepped like SB3 1.0 DummyVecEnv: auto-reset on done, terminal obs kept in infos."""

    def __init__(self, spec, n_envs: int, make_physics):
        self.spec, self.num_envs = spec, n_envs
        self.envs = [OracleMonitor(OracleMimicEnv(spec, make_physics())) for _ in range(n_envs)]

    def reset(self, inj_istep=None, inj_pos=None):
        return np.stack([e.env.reset(None if inj_istep is None else inj_istep[i], None if inj_pos is None else inj_pos[i])
                         for i, e in enumerate(self.envs)])

    def step(self, actions, inj_istep=None, inj_pos=None):
        obs, rews, dones, infos = [], [], [], []
        for i, e in enumerate(self.envs):
            o, r, d, info = e.step(actions[i])
            if d:
                info = dict(info)
                info["terminal_observation"] = o
                o = e.env.reset(None if inj_istep is None else inj_istep[i], None if inj_pos is None else inj_pos[i])
            obs.append(o); rews.append(r); dones.append(d); infos.append(info)
        return np.stack(obs), np.array(rews, dtype=np.float64), np.array(dones), infos


class RunningMeanStd:
    """SB3 1.0 common/running_mean_std.py (Chan parallel update, initial count epsilon=1e-4)."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr):
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class OracleVecNormalize:
    """SB3 1.0 VecNormalize(venv, norm_obs=True, norm_reward=True, clip 10/10, gamma=0.99, eps=1e-8) as built at
    reference drloco/common/utils.py:130-132 (note gamma=0.99 default, not hypers.gamma — Q16)."""

    def __init__(self, venv, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8, training=True):
        self.venv = venv
        self.obs_rms = RunningMeanStd(shape=(venv.spec.obs_dim,))
        self.ret_rms = RunningMeanStd(shape=())
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon
        self.norm_obs, self.norm_reward, self.training = norm_obs, norm_reward, training
        self.ret = np.zeros(venv.num_envs)

    def normalize_obs(self, obs):
        if self.norm_obs:
            obs = np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs,
                          self.clip_obs)
        return obs

    def normalize_reward(self, rew):
        if self.norm_reward:
            rew = np.clip(rew / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        return rew

    def step(self, actions, **kw):
        obs, rews, dones, infos = self.venv.step(actions, **kw)
        if self.training:
            self.obs_rms.update(obs)
        nobs = self.normalize_obs(obs)
        if self.training:
            self.ret = self.ret * self.gamma + rews
            self.ret_rms.update(self.ret)
        nrew = self.normalize_reward(rews)
        self.ret[dones] = 0
        return nobs, nrew, dones, infos

    def reset(self, reset_update="sb3-1.0", **kw):
        """SB3 1.0 VecNormalize.reset (restated from memory; SB3 is absent here): ``self.ret = zeros``, then
        ``self._update_reward(self.ret)`` when training - the observation statistics are NOT updated at reset in that
        release.  reset_update="obs" is the later behaviour (obs_rms.update(obs) as well)."""
        obs = self.venv.reset(**kw)
        self.ret = np.zeros(self.venv.num_envs)
        if self.training:
            self.ret_rms.update(self.ret)
            if reset_update == "obs":
                self.obs_rms.update(obs)
        return self.normalize_obs(obs)
