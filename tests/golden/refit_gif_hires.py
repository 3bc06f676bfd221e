"""Refine the FITTED initial states of tests/golden/gif_episodes.npz against the hi-res half of the recordings.

Object2-circles.gif / Object2-cubes.gif were recorded with forced shapes, i.e. not from the seeded reset, so their initial
poses are unknown; fit_gif_episodes.py searched for poses that reproduce the LCD half (0.31 m per pixel).  The left half
of the same gif frames shows the episode at 0.039 m per pixel (make_gif_hires.py); this script continues the search with
the objective "differing RGB pixels between the oracle replay drawn by boxlcd_b200/rgb_render.py and the recording, over
all frames" (random local search, shrinking step).  Result committed on 2026-10-18:
  Object2-circles: ZERO differing pixels on all 50 hi-res frames (was: up to 3 px off) -> stored;
  Object2-cubes:   no pose found beyond the LCD fit (11 exact frames, within 1 hi-res px for 41) -> unchanged.

Run in the build container only:   python tests/golden/refit_gif_hires.py Object2-circles
"""
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import boxlcd_b200 as blcd  # noqa: E402
from boxlcd_b200 import rgb_render  # noqa: E402
from oracle import oracle  # noqa: E402

VARIANT = {'Object2-circles': 0b00, 'Object2-cubes': 0b11}


def score(rec, cands, var):
  """cands [n, B, 3] -> (sum over frames of min(differing pixels, 200) * 0.93^t, leading exact frames, raw pixel count)"""
  env = blcd.env_map['Object2']()
  sp = env.layout.spec
  n = len(cands)
  bodies = np.zeros((n, sp.n_bodies, 6), np.float32)
  bodies[:, :, :3] = cands
  ow = oracle.OracleWorlds(sp, n, threads=os.cpu_count() or 1)
  ow.set_bodies(bodies, np.full(n, var, np.uint32))
  shapes = rgb_render.body_shapes(sp, var)
  tot, raw, prefix, alive, bad = np.zeros(n), np.zeros(n, int), np.zeros(n, int), np.ones(n, bool), np.zeros(n, int)
  for t in range(len(rec)):
    ow.step(np.zeros((n, sp.act_size), np.float32))
    poses, _ = ow.get_poses()
    for i in range(n):
      if bad[i] > 6:           # hopeless candidate: stop drawing it
        tot[i] += 200 * 0.93 ** t
        continue
      nd = int((rgb_render.render_rgb(shapes, poses[i], env.WIDTH, rec.shape[2], rec.shape[1]) != rec[t]).any(-1).sum())
      raw[i] += nd
      tot[i] += min(nd, 200) * 0.93 ** t
      bad[i] += nd > 150
      alive[i] &= nd == 0
      prefix[i] += alive[i]
  return tot, prefix, raw


def main(name, budget_s=1500):
  path = os.path.join(HERE, 'gif_episodes.npz')
  ep = dict(np.load(path))
  hi = np.load(os.path.join(HERE, 'gif_hires.npz'))
  rec = hi['palette'][hi[f'{name}_hi']]
  var = VARIANT[name]
  best = ep[f'{name}_init'].astype(np.float64)
  s, p, r = score(rec, best[None], var)
  bs, bp, br = s[0], p[0], r[0]
  print(name, 'start: score %.1f, exact prefix %d, differing pixels %d' % (bs, bp, br))
  rng = np.random.RandomState(0)
  step, t0 = 2e-2, time.time()
  dims = [(b, k) for b in range(best.shape[0]) for k in range(3 if var else 2)]   # a circle's angle never shows
  while step > 2e-6 and time.time() - t0 < budget_s and br > 0:
    cands = np.repeat(best[None], 96, 0)
    for i in range(1, 96):
      for (b, k) in dims:
        if rng.rand() < 0.5:
          cands[i, b, k] += rng.normal() * step * (3 if k == 2 else 1)
    s, p, r = score(rec, cands, var)
    j = int(np.argmin(s))
    if s[j] < bs:
      bs, bp, br, best = s[j], p[j], r[j], cands[j].copy()
      print('  step %.1e -> score %.1f, exact prefix %d, differing pixels %d' % (step, bs, bp, br), flush=True)
    else:
      step *= 0.6
  if br < int(score(rec, ep[f'{name}_init'].astype(np.float64)[None], var)[2][0]):
    ep[f'{name}_init'] = best.astype(np.float32)
    np.savez_compressed(path, **ep)
    print(name, 'stored', best.astype(np.float32).tolist())


if __name__ == '__main__':
  main(sys.argv[1])
