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


def oracle_sensitivity(env, bodies, variants, act, copies=48, seed=1, max_nudges=4, max_ulps=1):
  """How far the ORACLE's own single-step result moves when the inputs of a world are nudged by one fp32 ulp: for every
  world [m] run `copies` oracle copies whose initial state differs from the original in 1-4 randomly chosen components by
  one ulp, and return the largest relative deviation (position / angle) of any copy from the unnudged oracle.  Ordinary
  worlds give ~1e-7; a world that sits on a discrete decision boundary (a contact appearing or not, a clip point changing
  identity, a limit engaging, the block solver switching case, a TOI event) gives the size of the jump between the two
  outcomes."""
  from oracle import oracle
  import os
  sp = env.layout.spec
  m = len(bodies)
  rng = np.random.RandomState(seed)
  b2 = np.repeat(bodies, copies, 0).copy()
  for i in range(len(b2)):
    if i % copies == 0:
      continue
    for _ in range(rng.randint(1, max_nudges + 1)):
      b, k = rng.randint(sp.n_bodies), rng.randint(6)
      up = rng.rand() < .5
      for _ in range(rng.randint(1, max_ulps + 1)):
        v = b2[i, b, k]
        b2[i, b, k] = np.nextafter(v, np.float32(v + (1 if up else -1)), dtype=np.float32)
  ow = oracle.OracleWorlds(sp, len(b2), threads=os.cpu_count() or 1)
  ow.set_bodies(b2, None if variants is None else np.repeat(variants, copies))
  ow.step(np.repeat(act, copies, 0))
  o = ow.get_bodies().reshape(m, copies, sp.n_bodies, 6)
  return rel_err(o[:, 1:, :, :3], o[:, :1, :, :3]).max((1, 2, 3))


def rel_err(a, b, floor=1.0):
  return np.abs(a - b) / np.maximum(np.abs(b), floor)
