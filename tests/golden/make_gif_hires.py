"""Store the HI-RES half of the reference's recorded episodes (/root/reference/assets/envs/*.gif).

Every gif frame is `[lcd_render(8W, 8H, 'RGB') | 1 px | LCD frame x8]` (world_env.py:525-531, recorder
research/scripts/evaluations/demo_imgs.py:59-72).  tests/golden/gif_episodes.npz keeps the LCD half (0.31 m per pixel);
this script keeps the left half -- the same pybox2d episode at 8x the resolution (0.039 m per pixel), drawn by the
author's Pillow -- as palette indices in tests/golden/gif_hires.npz:
   <name>_hi   uint8 [T, 8H, 8W]   index into `palette`
   palette     uint8 [K, 3]        colours that occur (background 254, robot fill / outline, object fill / outline)
tests/test_gif_hires.py replays the episodes (initial states + actions of gif_episodes.npz) through the oracle and the
CUDA path, draws the same colour view from their body transforms and compares pixel by pixel.

Run in the build container only (needs /root/reference and PIL):   python tests/golden/make_gif_hires.py
"""
import os
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
GIFS = ['Bounce', 'Dropbox', 'Bounce2', 'Object2', 'Object2-circles', 'Object2-cubes', 'Urchin', 'Luxo', 'UrchinCube', 'UrchinBall', 'LuxoBall', 'LuxoCube']


def gif_frames(path):
  im = Image.open(path)
  frames = []
  try:
    while True:
      frames.append(np.array(im.convert('RGB')))
      im.seek(im.tell() + 1)
  except EOFError:
    pass
  return np.array(frames)


def main():
  out, palette = {}, []
  for name in GIFS:
    f = gif_frames(f'/root/reference/assets/envs/{name}.gif')
    w8 = f.shape[2] // 2      # the window is 2 x 8W wide: the picture's last LCD column is cut off (viewer.py:7)
    hi = f[:, :, :w8]
    assert (f[:, :, w8] == 0).all(), 'separator column'
    cols = np.unique(hi.reshape(-1, 3), axis=0)
    for c in cols:
      if tuple(c) not in palette:
        palette.append(tuple(c))
    idx = np.zeros(hi.shape[:3], np.uint8)
    for c in cols:
      idx[(hi == c).all(-1)] = palette.index(tuple(c))
    out[f'{name}_hi'] = idx
    print(name, idx.shape, 'colours', [tuple(int(x) for x in c) for c in cols])
  out['palette'] = np.asarray(palette, np.uint8)
  np.savez_compressed(os.path.join(HERE, 'gif_hires.npz'), **out)


if __name__ == '__main__':
  main()
