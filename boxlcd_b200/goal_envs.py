"""Batched, device-resident versions of the reference's goal-conditioned wrappers (research/wrappers/body_goal.py:15-103,
cube_goal.py:7-89): a goal is a second sampled state of the same world; rewards and `done` are computed from the
observation tensors with torch ops, so a GPU-resident RL loop never leaves the device.  Host glue over VecWorldEnv: the
simulation itself stays in libboxlcd_b200.

Reference semantics kept: BodyGoalEnv -- goal = obs of a fresh reset; state reward = -mean |goal - proprio| over the
`x:p` / `y:p` entries (or -0.05 + 10 * (last_delta - delta) with G.diff_delt), +1 and done below G.goal_thresh; LCD reward
= -1 + (ink pixels shared with the goal / ink pixels), 0 and done above 0.70.  CubeGoalEnv -- goal = state after 10
zero-action steps of a fresh reset; delta over the object's x:p / y:p; +1 and done below 0.05.  Rewards are scaled by
G.rew_scale; `done` also includes the env's own timeout."""
import re
import torch


def _idx(keys, pattern):
  return [i for i, k in enumerate(keys) if re.match(pattern, k) is not None]


class _GoalBase:
  def __init__(self, vec, G):
    self.vec, self.env, self.G = vec, vec.env, G
    self.n = vec.n
    self.action_space = vec.action_space
    self.goal, self.last = None, None

  def _snapshot(self, obs):
    return {k: v.clone() for k, v in obs.items() if k in ('full_state', 'proprio', 'lcd_bits')}

  def _attach(self, obs):
    out = dict(obs)
    out['goal:lcd_bits'], out['goal:proprio'], out['goal:full_state'] = self.goal['lcd_bits'], self.goal['proprio'], self.goal['full_state']
    return out

  def _finish(self, obs, rew, goal_done):
    done = obs['done'].bool() | goal_done
    rew = rew * float(getattr(self.G, 'rew_scale', 1.0))
    self.last = self._snapshot(obs)
    return self._attach(obs), rew, done, {'success': goal_done}

  def close(self):
    self.vec.close()


class VecBodyGoalEnv(_GoalBase):
  """research/wrappers/body_goal.py"""

  def __init__(self, vec, G):
    super().__init__(vec, G)
    self.xy_idx = torch.tensor(_idx(self.env.pobs_keys, '.*(x|y):p'), device=vec.device, dtype=torch.long)
    self.full = (1 << vec.W) - 1 if vec.W < 32 else 0xFFFFFFFF

  def reset(self):
    self.vec.reset_dev()
    self.goal = self._snapshot(self.vec.observe_dev())
    self.vec.reset_dev()
    obs = self.vec.observe_dev()
    self.last = self._snapshot(obs)
    return self._attach(obs)

  def _delta(self, proprio):
    return (self.goal['proprio'] - proprio).abs()[:, self.xy_idx].mean(1)

  def _ink_counts(self, bits):
    return (~self.vec.unpack_lcd(bits)).to(torch.int32)   # [N, H, W], 1 = body pixel

  def comp_rew_done(self, obs):
    if getattr(self.G, 'state_rew', 1):
      delta = self._delta(obs['proprio'])
      if getattr(self.G, 'diff_delt', 0):
        rew = -0.05 + 10 * (self._delta(self.last['proprio']) - delta)
      else:
        rew = -delta
      hit = delta < float(getattr(self.G, 'goal_thresh', 0.01))
      return rew + hit.float(), hit
    ink, gink = self._ink_counts(obs['lcd_bits']), self._ink_counts(self.goal['lcd_bits'])
    n_ink = ink.sum((1, 2)).float()
    similarity = (ink & gink).sum((1, 2)).float() / n_ink          # == mean(lcd==0 & lcd==goal) / mean(lcd==0)
    hit = similarity > 0.70
    return torch.where(hit, torch.zeros_like(similarity), -1 + similarity), hit

  def step(self, actions=None):
    obs, _ = self.vec.step_dev(actions, observe=True)
    obs = {k: v for k, v in obs.items()}
    rew, hit = self.comp_rew_done(obs)
    return self._finish(obs, rew, hit)


class VecCubeGoalEnv(_GoalBase):
  """research/wrappers/cube_goal.py"""

  def __init__(self, vec, G):
    super().__init__(vec, G)
    self.obj_idx = torch.tensor(_idx(self.env.obs_keys, 'object.*(x|y):p'), device=vec.device, dtype=torch.long)

  def reset(self):
    self.vec.reset_dev()
    zero = torch.zeros((self.n, self.vec.A), device=self.vec.device)
    for _ in range(10):
      obs, _ = self.vec.step_dev(zero, observe=True)
    self.goal = self._snapshot(obs)
    self.vec.reset_dev()
    obs = self.vec.observe_dev()
    self.last = self._snapshot(obs)
    out = self._attach(obs)
    out['goal:object'] = self.goal['full_state'][:, self.obj_idx]
    return out

  def _delta(self, full_state):
    return (self.goal['full_state'][:, self.obj_idx] - full_state[:, self.obj_idx]).abs().mean(1)

  def comp_rew_done(self, obs):
    delta = self._delta(obs['full_state'])
    if getattr(self.G, 'diff_delt', 0):
      rew = -0.05 + 10 * (self._delta(self.last['full_state']) - delta)
    else:
      rew = -delta
    hit = delta < 0.05
    return rew + hit.float(), hit

  def step(self, actions=None):
    obs, _ = self.vec.step_dev(actions, observe=True)
    obs = {k: v for k, v in obs.items()}
    rew, hit = self.comp_rew_done(obs)
    out, rew, done, info = self._finish(obs, rew, hit)
    out['goal:object'] = self.goal['full_state'][:, self.obj_idx]
    return out, rew, done, info
