"""lcd_mode='RGB' (world_env.py:473-474, 481-483, 509-511) and the human-view picture (world_env.py:525-531):
boxlcd_b200/rgb_render.py against frames of the UNMODIFIED reference renderer (tests/golden/make_rgb_golden.py), at the
LCD size and at the viewer's x8 size, bit for bit."""
import os
import numpy as np
import pytest
import boxlcd_b200 as blcd
from boxlcd_b200 import rgb_render

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'rgb_golden.npz')
ENVS = ['Dropbox', 'Bounce2', 'Object2', 'Urchin', 'Luxo', 'UrchinCube', 'LuxoCube', 'UrchinBall', 'LuxoBall', 'Crab']


@pytest.fixture(scope='module')
def gold():
  import PIL
  g = np.load(GOLD)
  if str(g['pillow_version']) != PIL.__version__:
    pytest.skip(f'golden colour frames were drawn by Pillow {g["pillow_version"]}, this is {PIL.__version__}')
  return g


@pytest.mark.parametrize('name', ENVS)
def test_rgb_frames_bit_exact_vs_reference_golden(gold, name):
  env = blcd.env_map[name]()
  sp = env.layout.spec
  world_w, W, H = [int(x) for x in gold[f'{name}_meta']]
  assert (world_w, W, H) == (env.WIDTH, sp.lcd_w, sp.lcd_h)
  poses, variant = gold[f'{name}_poses'], gold[f'{name}_variant']
  for key, scale in (('rgb', 1), ('rgb8', 8)):
    want = gold[f'{name}_{key}']
    for i in range(len(poses)):
      got = rgb_render.render_rgb(rgb_render.body_shapes(sp, variant[i]), poses[i], world_w, W * scale, H * scale)
      assert got.dtype == np.uint8 and got.shape == want[i].shape
      assert (got == want[i]).all(), f'{name} frame {i} x{scale}: {int((got != want[i]).any(-1).sum())} pixels differ'
  assert len(np.unique(gold[f'{name}_rgb8'].reshape(-1, 3), axis=0)) >= 3   # background, fill, outline


def test_human_frame_layout():
  hi = np.full((128, 256, 3), 7, np.uint8)
  lcd = np.zeros((16, 32), bool)
  lcd[0, 1] = True
  img = rgb_render.human_frame(hi, lcd)
  assert img.shape == (128, 256 + 1 + 256, 3) and img.dtype == np.uint8
  assert (img[:, :256] == 7).all() and (img[:, 256] == 0).all()
  assert (img[:8, 257 + 8:257 + 16] == 255).all() and (img[:8, 257:257 + 8] == 0).all()
  rgb = np.full((16, 32, 3), 9, np.uint8)
  assert (rgb_render.human_frame(hi, rgb)[:, 257:] == 9).all()
