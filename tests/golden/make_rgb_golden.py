"""Generate tests/golden/rgb_golden.npz: colour frames produced by the UNMODIFIED reference
`WorldEnv.lcd_render(width, height, lcd_mode='RGB')` (/root/reference/boxLCD/world_env.py:460-512) on this container's
Pillow for random body poses, at the env's LCD size and at the human viewer's x8 size (world_env.py:525).

Run in the build container only (needs /root/reference):   python tests/golden/make_rgb_golden.py
The fixture pins boxlcd_b200/rgb_render.py (tests/test_rgb_render.py).

Per env: poses [n, B, 4] f32 (x, y, sin, cos; draw order), variant [n] (bit b = body b drawn as its second shape),
rgb [n, H, W, 3] uint8, rgb8 [n, 8H, 8W, 3] uint8, meta = (WIDTH, W, H).
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from make_lcd_golden import random_poses  # noqa: E402

ENVS = ['Dropbox', 'Bounce2', 'Object2', 'Urchin', 'Luxo', 'UrchinCube', 'LuxoCube', 'UrchinBall', 'LuxoBall', 'Crab']


def render(env, poses, w, h):
  for (name, body), (px, py, s, c) in zip(env.dynbodies.items(), poses):
    body.position = (px, py)
    body._sc = (np.float32(s), np.float32(c))
  return env.lcd_render(w, h, lcd_mode='RGB')


def main(n_per_env=24, seed=0):
  boxLCD = ref_harness.ref_envs()
  out = {}
  rng = np.random.RandomState(seed)
  np.random.seed(seed)
  for name in ENVS:
    env = boxLCD.env_map[name]()
    env.seed(seed)
    W, H = int(env.G.lcd_base * env.G.wh_ratio), env.G.lcd_base
    P, V, lo, hi = [], [], [], []
    for i in range(n_per_env):
      if i % 4 == 0:
        env.reset()
      bodies = list(env.dynbodies.values())
      poses = random_poses(rng, env, len(bodies), i % 3)
      V.append(sum((0 if isinstance(b.fixtures[0].shape, ref_harness.circleShape) else 1) << k for k, b in enumerate(bodies)
                   if name.startswith('Object') ))
      P.append(poses)
      lo.append(render(env, poses, W, H))
      hi.append(render(env, poses, 8 * W, 8 * H))
    out[f'{name}_poses'] = np.asarray(P, np.float32)
    out[f'{name}_variant'] = np.asarray(V, np.uint32)
    out[f'{name}_rgb'] = np.asarray(lo, np.uint8)
    out[f'{name}_rgb8'] = np.asarray(hi, np.uint8)
    out[f'{name}_meta'] = np.asarray([env.WIDTH, W, H], np.int32)
    print(name, out[f'{name}_rgb8'].shape, 'colours', len(np.unique(out[f'{name}_rgb8'].reshape(-1, 3), axis=0)))
  import PIL
  out['pillow_version'] = np.asarray(PIL.__version__)
  np.savez_compressed(os.path.join(HERE, 'rgb_golden.npz'), **out)


if __name__ == '__main__':
  main()
