"""Pin the physics oracle against REAL pybox2d output: the episodes the reference author recorded into
`/root/reference/assets/envs/*.gif` (each frame = [hi-res render | LCD frame upscaled 8x]).  For the passive envs (no
random actions) an episode is a deterministic function of the initial poses, so this script searches for initial poses
(bodies at rest, as after WorldEnv.reset) whose simulated LCD frames reproduce EVERY frame of the gif exactly, and stores
the gif's LCD frames plus one such initial state per gif in tests/golden/gif_episodes.npz.

Run in the build container only (needs /root/reference and PIL):   python tests/golden/fit_gif_episodes.py
Gif frame f shows the state after f + 1 env steps (the recorder rendered after stepping).
tests/test_gif_episodes.py replays the stored initial states through the oracle (and the CUDA path on a GPU).
"""
import os
import sys
import time
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import boxlcd_b200 as blcd  # noqa: E402
from oracle import oracle  # noqa: E402

GIFS = {  # gif -> (env, variant bitmask of the shapes shown, per-body search is over x, y[, angle])
    'Bounce': ('Bounce', 0), 'Dropbox': ('Dropbox', 0), 'Bounce2': ('Bounce2', 0),
    'Object2-circles': ('Object2', 0b00), 'Object2-cubes': ('Object2', 0b11), 'Object2': ('Object2', None),
}


def gif_lcd(path):
  im = Image.open(path)
  frames = []
  try:
    while True:
      frames.append(np.array(im.convert('RGB')))
      im.seek(im.tell() + 1)
  except EOFError:
    pass
  f = np.array(frames)
  w8 = (f.shape[2] - 1) // 2
  return f[:, 4::8, w8 + 1 + 4::8, 0] > 127     # True = background, like the reference's lcd array


def run(spec, cands, variants, lcd, T):
  """cands [n, B, 3] (x, y, angle) at rest -> number of leading gif frames reproduced exactly"""
  n, B = cands.shape[:2]
  bodies = np.zeros((n, B, 6), np.float32)
  bodies[..., :3] = cands
  ow = oracle.OracleWorlds(spec, n, threads=os.cpu_count() or 1)
  ow.set_bodies(bodies, variants)
  match = np.ones(n, bool)
  prefix = np.zeros(n, int)
  W = lcd.shape[2]
  for t in range(T):
    ow.step(np.zeros((n, spec.act_size), np.float32))
    ok = (oracle.unpack_bits(ow.observe()['lcd_bits'], W) == lcd[t]).all((1, 2))
    prefix += match & ok
    match &= ok
    if not match.any():
      break
  return prefix


def fit(name, env_name, variant, lcd, rng, budget_s=600):
  env = blcd.env_map[env_name]()
  spec = env.layout.spec
  B, T = spec.n_bodies, lcd.shape[0]
  size = [spec.bodies[b].extent for b in range(B)]
  variants_opts = [variant] if variant is not None else [0b01, 0b10]
  t0 = time.time()
  for var in variants_opts:
    is_box = [bool((var >> b) & 1) if spec.bodies[b].n_variants > 1 else spec.bodies[b].shape[0].kind != 0 for b in range(B)]
    def sample(n):
      c = np.zeros((n, B, 3), np.float32)
      for b in range(B):
        c[:, b, 0] = rng.uniform(size[b], env.WIDTH - size[b], n)
        c[:, b, 1] = rng.uniform(size[b], env.HEIGHT - size[b], n)
        c[:, b, 2] = rng.uniform(-np.pi / 4, np.pi / 4, n) if is_box[b] else 0.0    # a square repeats every 90 degrees
      return c
    keep = np.zeros((0, B, 3), np.float32)
    rounds = 0
    while len(keep) < 8 and time.time() - t0 < budget_s / 2 and rounds < 40:      # frames 0..1 filter on blind samples
      c = sample(400000)
      p = run(spec, c, np.full(len(c), var, np.uint32), lcd, 2)
      keep = np.concatenate([keep, c[p >= 2]])
      rounds += 1
    print(f'{name}: variant {var:02b}: {len(keep)} candidates reproduce the first two frames ({time.time() - t0:.0f} s)')
    if len(keep) == 0:
      continue
    best_c, best = None, 0
    for it, scale in enumerate([0.05, 0.03, 0.015, 0.008, 0.004, 0.002, 0.001, 0.0005, 0.00025]):
      k = keep[rng.randint(len(keep), size=150000)]
      noise = rng.normal(size=k.shape) * np.array([scale, scale, 2 * scale])
      noise[..., 2] *= np.array(is_box)[None, :]
      c = np.concatenate([keep[:20000], (k + noise).astype(np.float32)])
      p = run(spec, c, np.full(len(c), var, np.uint32), lcd, T)
      best = int(p.max())
      keep = c[p >= max(best - 1, 2)]
      best_c = c[p == best]
      print(f'  refine {it}: best prefix {best}/{T}, {len(best_c)} candidates ({time.time() - t0:.0f} s)')
      if best == T and len(best_c) >= 50:
        break
      if time.time() - t0 > budget_s:
        break
    if best == T:
      # the matching candidate closest to the cluster centre: largest margin against last-bit differences
      centre = np.median(best_c, 0)
      pick = best_c[np.argmin(np.abs(best_c - centre).sum((1, 2)))]
      return pick, var, T
    print(f'{name}: variant {var:02b}: best prefix {best}/{T}')
  return None, None, 0


def main():
  rng = np.random.RandomState(0)
  out = {}
  only = sys.argv[1:] or list(GIFS)
  path = os.path.join(HERE, 'gif_episodes.npz')
  if os.path.exists(path):
    out = dict(np.load(path))
  for name in only:
    env_name, variant = GIFS[name]
    lcd = gif_lcd(f'/root/reference/assets/envs/{name}.gif')
    init, var, T = fit(name, env_name, variant, lcd, rng)
    out[f'{name}_lcd'] = np.packbits(lcd, axis=2)
    out[f'{name}_shape'] = np.array(lcd.shape)
    if init is not None:
      out[f'{name}_init'] = init
      out[f'{name}_variant'] = np.array(var, np.uint32)
      print(f'{name}: all {T} frames reproduced from initial poses\n{init}')
    else:
      print(f'{name}: NO initial state found that reproduces the whole gif')
  np.savez_compressed(path, **out)


if __name__ == '__main__':
  main()
