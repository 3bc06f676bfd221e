"""Generate tests/golden/lcd_golden.npz: LCD frames produced by the UNMODIFIED reference `WorldEnv.lcd_render`
(/root/reference/boxLCD/world_env.py:460-512) on this container's Pillow, for random body poses.

Run in the build container only (needs /root/reference):   python tests/golden/make_lcd_golden.py
The fixture pins oracle/lcd_oracle.c (tests/test_lcd_oracle.py) and, through it, the CUDA rasterizer.

Per env the file holds: poses [n,B,4] f32 (x, y, sin, cos of each dynamic body's b2Transform, draw order), per-world
shape tables (kind [n,B], nvert [n,B], radius [n,B], verts [n,B,8,2]) and the reference frame bit-packed as
bits [n,H] uint32 (bit x = pixel x, 1 = background; [n,H,2] for the 64-px-wide Crab / SpiderCube frames), plus meta = (WIDTH, W, H).
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

F = np.float32
ENVS = ['Dropbox', 'Bounce2', 'Object2', 'Urchin', 'Luxo', 'UrchinCube', 'LuxoCube', 'UrchinBall', 'LuxoBall', 'Crab', 'SpiderCube']


def pack_bits(lcd):
  """[H, W] bool -> [H] uint32, or [H, ceil(W / 32)] (word k = pixels 32k .. 32k+31) for frames wider than 32 px"""
  H, W = lcd.shape
  if W <= 32:
    return (lcd.astype(np.uint64) << np.arange(W, dtype=np.uint64)[None]).sum(1).astype(np.uint32)
  lw = (W + 31) // 32
  pad = np.zeros((H, 32 * lw), np.uint64)
  pad[:, :W] = lcd
  return (pad.reshape(H, lw, 32) << np.arange(32, dtype=np.uint64)).sum(2).astype(np.uint32)


def shape_row(body):
  shape = body.fixtures[0].shape
  verts = np.zeros((8, 2), F)
  if isinstance(shape, ref_harness.circleShape):
    return 0, 0, F(shape.radius), verts
  vs = np.asarray(shape.vertices, F)
  verts[:len(vs)] = vs
  return 1, len(vs), F(0), verts


def random_poses(rng, env, n_bodies, mode):
  W, Hh = env.WIDTH, env.HEIGHT
  poses = np.zeros((n_bodies, 4), F)
  for b in range(n_bodies):
    if mode == 0:      # anywhere inside the arena
      x, y = rng.uniform(0, W), rng.uniform(0, Hh)
    elif mode == 1:    # hugging a wall / corner, partly off-canvas on the high side
      x = rng.choice([rng.uniform(0, 0.8), rng.uniform(W - 0.8, W + 0.3), rng.uniform(0, W)])
      y = rng.choice([rng.uniform(0, 0.8), rng.uniform(Hh - 0.8, Hh + 0.3), rng.uniform(0, Hh)])
    else:              # resting on the floor, axis aligned or nearly so
      x, y = rng.uniform(0.3, W - 0.3), rng.uniform(0.1, 1.5)
    ang = rng.uniform(-np.pi, np.pi)
    if mode == 2 and rng.uniform() < 0.7:
      ang = rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2]) + rng.choice([0.0, rng.normal() * 1e-3])
    a32 = F(ang)
    poses[b] = (F(x), F(y), F(np.sin(np.float64(a32))), F(np.cos(np.float64(a32))))
  return poses


def main(n_per_env=600, seed=0):
  boxLCD = ref_harness.ref_envs()
  out = {}
  rng = np.random.RandomState(seed)
  np.random.seed(seed)
  for name in ENVS:
    env = boxLCD.env_map[name]()
    env.seed(seed)
    all_poses, kinds, nverts, radii, verts, bits = [], [], [], [], [], []
    for i in range(n_per_env):
      if i % 50 == 0:
        env.reset()   # re-draws Object2's random shapes (global np.random, world_env.py:274)
      bodies = list(env.dynbodies.values())
      poses = random_poses(rng, env, len(bodies), i % 3)
      lcd = ref_harness.render_poses(env, poses)
      rows = [shape_row(b) for b in bodies]
      all_poses.append(poses)
      kinds.append([r[0] for r in rows]); nverts.append([r[1] for r in rows])
      radii.append([r[2] for r in rows]); verts.append([r[3] for r in rows])
      bits.append(pack_bits(np.asarray(lcd, bool)))
    out[f'{name}_poses'] = np.asarray(all_poses, F)
    out[f'{name}_kind'] = np.asarray(kinds, np.int32)
    out[f'{name}_nvert'] = np.asarray(nverts, np.int32)
    out[f'{name}_radius'] = np.asarray(radii, F)
    out[f'{name}_verts'] = np.asarray(verts, F)
    out[f'{name}_bits'] = np.asarray(bits, np.uint32)
    out[f'{name}_meta'] = np.asarray([env.WIDTH, int(env.G.lcd_base * env.G.wh_ratio), env.G.lcd_base], np.int32)
    print(name, out[f'{name}_poses'].shape, 'ink mean', float(np.mean([bin(int(b)).count('0') for b in np.asarray(bits).ravel()[:64]])))
  # the x8 view the reference's human renderer asks for (world_env.py:525: lcd_render(width * 8, height * 8)), mode '1'
  for name, n8 in [('Urchin', 60), ('LuxoCube', 60)]:
    env = boxLCD.env_map[name]()
    env.seed(seed)
    env.reset()
    bodies = list(env.dynbodies.values())
    W8, H8 = int(env.G.lcd_base * env.G.wh_ratio) * 8, env.G.lcd_base * 8
    poses = [random_poses(rng, env, len(bodies), i % 3) for i in range(n8)]
    out[f'{name}_x8_poses'] = np.asarray(poses, F)
    out[f'{name}_x8_bits'] = np.asarray([pack_bits(np.asarray(ref_harness.render_poses(env, p, W8, H8), bool)) for p in poses], np.uint32)
    out[f'{name}_x8_meta'] = np.asarray([env.WIDTH, W8, H8], np.int32)
    print(name, 'x8', out[f'{name}_x8_bits'].shape)
  import PIL
  out['pillow_version'] = np.asarray(PIL.__version__)
  np.savez_compressed(os.path.join(HERE, 'lcd_golden.npz'), **out)


if __name__ == '__main__':
  main()
