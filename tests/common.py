"""shared helpers for the parity tests"""
import numpy as np
import boxlcd_b200 as blcd

ENVS_CORE = ['Dropbox', 'Bounce2', 'Object2', 'Urchin', 'Luxo', 'LuxoCube', 'UrchinBall']


def make_env(name, **G):
  return blcd.env_map[name](G)


def random_bodies(env, n, rng, speed=3.0):
  """random but plausible fresh-world states [n, B, 6] for the single-step parity protocol: take the reference reset
  distribution's geometry (bodies inside the arena, robot parts hinged together) by sampling oracle resets, then add
  random velocities."""
  from oracle import oracle
  ow = oracle.OracleWorlds(env.layout.spec, n, seed=int(rng.randint(1 << 30)))
  ow.reset()
  b = ow.get_bodies()
  _, variants = ow.get_poses()
  b[..., 3:5] = rng.uniform(-speed, speed, b[..., 3:5].shape)
  b[..., 5] = rng.uniform(-speed, speed, b[..., 5].shape)
  # drop a share of the worlds onto the floor so that contacts are exercised from step one
  low = rng.uniform(size=n) < 0.5
  shift = b[:, :, 1].min(1) - rng.uniform(0.05, 0.6, n)
  b[low, :, 1] -= np.maximum(shift[low], 0)[:, None] * 0.9
  return b.astype(np.float32), variants


def rel_err(a, b, floor=1.0):
  return np.abs(a - b) / np.maximum(np.abs(b), floor)
