"""Run on ANY machine that has the real reference stack (pip install Box2D==2.3.10 gym==0.17.3 Pillow numpy, plus the
boxLCD repo on PYTHONPATH) to produce single-step physics golden vectors this repo cannot generate itself (pybox2d is not installable
in the build image; the physics is pinned there through the reference's recorded episodes instead, tests/test_gif_episodes.py).

  python tools/dump_pybox2d_golden.py --out tests/golden/pybox2d_steps.npz [--n 256]

Protocol (last paragraph): for each env, sample a state with the reference reset, read it back as
full_state s0, then `env.reset(full_state=s0); obs1 = env.step(a)`: a fresh b2World, no hidden warm-start history.
Stored per env: s0 [n, S], action [n, A], s1 [n, S], lcd1 [n, H, W], raw body states before/after [n, B, 6]
(x, y, angle, vx, vy, omega), and Box2D / Pillow versions.  Replaying robots from such a file needs care: a joint's
reference angle is fixed by the poses reset() SAMPLED before full_state was applied (pybox2d sets it when the joint is
defined), so it can differ by a multiple of 2 pi from what b0 alone implies.
"""
import argparse
import numpy as np


def body_state(env):
  out = []
  for b in env.dynbodies.values():
    out.append([b.position[0], b.position[1], b.angle, b.linearVelocity[0], b.linearVelocity[1], b.angularVelocity])
  return np.asarray(out, np.float32)


def main():
  p = argparse.ArgumentParser()
  p.add_argument('--out', default='tests/golden/pybox2d_steps.npz')
  p.add_argument('--n', type=int, default=256)
  a = p.parse_args()
  import Box2D
  import PIL
  from boxLCD import env_map
  out = {'box2d_version': np.asarray(getattr(Box2D, '__version__', 'unknown')), 'pillow_version': np.asarray(PIL.__version__)}
  for name in ['Dropbox', 'Bounce', 'Bounce2', 'Urchin', 'Luxo', 'UrchinCube', 'LuxoCube', 'UrchinBall', 'LuxoBall']:
    env = env_map[name]()
    env.seed(0)
    env.action_space.seed(0)
    rows = {k: [] for k in ('s0', 'action', 's1', 'lcd0', 'lcd1', 'b0', 'b1')}
    for i in range(a.n):
      s0 = env.reset()['full_state']
      obs0 = env.reset(full_state=s0)
      b0 = body_state(env)
      act = env.action_space.sample()
      obs1, _, _, _ = env.step(act)
      rows['s0'].append(obs0['full_state']); rows['action'].append(act); rows['s1'].append(obs1['full_state'])
      rows['lcd0'].append(obs0['lcd']); rows['lcd1'].append(obs1['lcd']); rows['b0'].append(b0); rows['b1'].append(body_state(env))
    for k, v in rows.items():
      out[f'{name}_{k}'] = np.asarray(v)
    print(name, 'done')
  np.savez_compressed(a.out, **out)


if __name__ == '__main__':
  main()
