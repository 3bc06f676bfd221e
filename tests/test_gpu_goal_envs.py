"""Batched goal wrappers vs a numpy restatement of the reference's per-env formulas
(research/wrappers/body_goal.py:58-88, cube_goal.py:64-86)."""
import numpy as np
import pytest
import torch
from types import SimpleNamespace
import boxlcd_b200 as blcd
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_body_goal_state_and_lcd_rewards():
  from boxlcd_b200.vec_env import VecWorldEnv
  from boxlcd_b200.goal_envs import VecBodyGoalEnv
  env = blcd.envs.Urchin()
  n = 256
  for G in (SimpleNamespace(state_rew=1, diff_delt=0, goal_thresh=0.05, rew_scale=2.0), SimpleNamespace(state_rew=1, diff_delt=1, goal_thresh=0.05, rew_scale=1.0),
            SimpleNamespace(state_rew=0, diff_delt=0, goal_thresh=0.05, rew_scale=1.0)):
    g = VecBodyGoalEnv(VecWorldEnv(env, n, seed=3), G)
    obs0 = g.reset()
    last_p = obs0['proprio'].cpu().numpy().copy()
    obs, rew, done, info = g.step(None)
    p, gp = obs['proprio'].cpu().numpy(), obs['goal:proprio'].cpu().numpy()
    lcd = oracle.unpack_bits(obs['lcd_bits'].cpu().numpy().view(np.uint32), 32)
    glcd = oracle.unpack_bits(obs['goal:lcd_bits'].cpu().numpy().view(np.uint32), 32)
    idxs = [i for i, k in enumerate(env.pobs_keys) if k.endswith('x:p') or k.endswith('y:p')]
    for w in range(0, n, 17):
      if G.state_rew:
        delta = np.abs(gp[w] - p[w])[idxs].mean()
        r = (-0.05 + 10 * (np.abs(gp[w] - last_p[w])[idxs].mean() - delta)) if G.diff_delt else -delta
        d = delta < G.goal_thresh
        r = r + 1.0 if d else r
      else:
        sim = np.logical_and(lcd[w] == 0, lcd[w] == glcd[w]).mean() / (lcd[w] == 0).mean()
        d = sim > 0.70
        r = 0 if d else -1 + sim
      assert rew[w].item() == pytest.approx(r * G.rew_scale, abs=2e-6) and bool(info['success'][w]) == bool(d)
    assert not (obs['goal:proprio'] == obs['proprio']).all()
    g.close()


def test_cube_goal_rewards():
  from boxlcd_b200.vec_env import VecWorldEnv
  from boxlcd_b200.goal_envs import VecCubeGoalEnv
  env = blcd.envs.LuxoCube()
  n = 128
  G = SimpleNamespace(diff_delt=1, rew_scale=1.0)
  g = VecCubeGoalEnv(VecWorldEnv(env, n, seed=5), G)
  obs0 = g.reset()
  assert obs0['goal:object'].shape == (n, 2)
  fs0 = obs0['full_state'].cpu().numpy().copy()
  obs, rew, done, info = g.step(torch.zeros((n, 3), device='cuda'))
  fs, gfs = obs['full_state'].cpu().numpy(), obs['goal:full_state'].cpu().numpy()
  idxs = [i for i, k in enumerate(env.obs_keys) if k.startswith('object') and (k.endswith('x:p') or k.endswith('y:p'))]
  for w in range(0, n, 9):
    delta = np.abs(gfs[w, idxs] - fs[w, idxs]).mean()
    last = np.abs(gfs[w, idxs] - fs0[w, idxs]).mean()
    r = -0.05 + 10 * (last - delta) + (1.0 if delta < 0.05 else 0.0)
    assert rew[w].item() == pytest.approx(r, abs=2e-6) and bool(info['success'][w]) == (delta < 0.05)
  # the goal is a settled state: after 10 zero-action steps the cube lies lower than where a fresh reset drops it from
  assert (gfs[:, idxs[1]] <= obs0['goal:full_state'].cpu().numpy()[:, idxs[1]] + 1e-6).all()
  g.close()
