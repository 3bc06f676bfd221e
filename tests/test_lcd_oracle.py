"""Pins oracle/lcd_oracle.c against the reference renderer: (i) the committed golden frames produced by the unmodified
`WorldEnv.lcd_render` (tests/golden/make_lcd_golden.py), (ii) in the build container, fresh random poses rendered live
by the reference under the stub harness."""
import os
import sys
import numpy as np
import pytest
from oracle import oracle

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'lcd_golden.npz')
ENVS = ['Dropbox', 'Bounce2', 'Object2', 'Urchin', 'Luxo', 'UrchinCube', 'LuxoCube', 'UrchinBall', 'LuxoBall', 'Crab', 'SpiderCube']


@pytest.fixture(scope='module')
def gold():
  return np.load(GOLD)


@pytest.mark.parametrize('env', ENVS)
def test_oracle_matches_reference_golden_frames(gold, env):
  shapes = oracle.make_shapes(gold[f'{env}_kind'], gold[f'{env}_nvert'], gold[f'{env}_radius'], gold[f'{env}_verts'])
  world_w, lcd_w, lcd_h = [int(x) for x in gold[f'{env}_meta']]
  bits = oracle.lcd_render(shapes, gold[f'{env}_poses'], world_w, lcd_w, lcd_h)
  assert bits.shape == gold[f'{env}_bits'].shape
  bad = np.nonzero((bits != gold[f'{env}_bits']).reshape(len(bits), -1).any(1))[0]
  assert len(bad) == 0, f'{env}: {len(bad)} / {len(bits)} frames differ, first {bad[:5]}'
  assert (~oracle.unpack_bits(bits, lcd_w)).any(), 'frames are empty'


@pytest.mark.parametrize('env', ['Urchin', 'LuxoCube'])
def test_oracle_matches_reference_at_the_human_view_size(gold, env):
  """lcd_render(width * 8, height * 8) (world_env.py:525), mode '1': 256 x 128 / 192 x 128 frames of the unmodified renderer"""
  import boxlcd_b200 as blcd
  sp = blcd.env_map[env]().layout.spec
  ow = oracle.OracleWorlds(sp, 1)
  ow.reset()
  world_w, w8, h8 = [int(x) for x in gold[f'{env}_x8_meta']]
  bits = oracle.lcd_render(ow.lcd_shapes(0), gold[f'{env}_x8_poses'], world_w, w8, h8)
  assert bits.shape == gold[f'{env}_x8_bits'].shape and (bits == gold[f'{env}_x8_bits']).all()
  assert (~oracle.unpack_bits(bits, w8)).sum() > 100 * len(bits)


@pytest.mark.reference
@pytest.mark.parametrize('env', ['Urchin', 'LuxoCube', 'UrchinBall', 'Object2', 'CrabCube'])
def test_oracle_matches_live_reference(env):
  sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
  import ref_harness
  import make_lcd_golden as mk
  boxLCD = ref_harness.ref_envs()
  e = boxLCD.env_map[env]()
  e.seed(123)
  rng = np.random.RandomState(123)
  np.random.seed(123)
  n = 1500
  poses, rows, ref = [], [], []
  for i in range(n):
    if i % 100 == 0:
      e.reset()
    bodies = list(e.dynbodies.values())
    p = mk.random_poses(rng, e, len(bodies), i % 3)
    ref.append(mk.pack_bits(np.asarray(ref_harness.render_poses(e, p), bool)))
    poses.append(p)
    rows.append([mk.shape_row(b) for b in bodies])
  shapes = oracle.make_shapes([[r[0] for r in row] for row in rows], [[r[1] for r in row] for row in rows],
                              [[r[2] for r in row] for row in rows], [[r[3] for r in row] for row in rows])
  W = int(e.G.lcd_base * e.G.wh_ratio)
  bits = oracle.lcd_render(shapes, np.asarray(poses, np.float32), e.WIDTH, W, e.G.lcd_base)
  assert (bits == np.asarray(ref)).all()


def test_unpack_bits_layout():
  bits = np.array([[0b101, 0xFFFFFFFF]], np.uint32)
  lcd = oracle.unpack_bits(bits, 4)
  assert lcd.shape == (1, 2, 4)
  assert lcd[0, 0].tolist() == [True, False, True, False] and lcd[0, 1].all()
