"""`lcd_render(width, height, lcd_mode='RGB')` and the human-view frame (boxLCD/world_env.py:460-512, 514-535).

SURVEY 8f-3: the colour view is the viewer's picture, not the dataset path, so it is drawn on the host with Pillow from
the body transforms the CUDA simulator holds (`blcd_get_poses`: position and (sin, cos) of every dynamic body exactly as
the rasterizer sees them).  The arithmetic mirrors the reference line by line:
  * polygons: `trans * v` in fp32 with separately rounded multiply / add (b2Mul on x86-64), then float64 `/ WIDTH * width`
    (world_env.py:502-505);
  * circles: float64 `(pos -/+ radius) / WIDTH * width` on the fp32 position (world_env.py:494-499);
  * fill / outline = `int(255 * (1 - c))` of color1 / color2 (world_env.py:482-483): robots (0.9, 0.4, 0.4) / (0.5, 0.3, 0.5)
    (world_env.py:201), objects (0.5, 0.4, 0.9) / (0.3, 0.3, 0.5) (world_env.py:303); background (1, 1, 1);
  * FLIP_TOP_BOTTOM, then `255 - image` (world_env.py:506-511).
Pinned bit-for-bit against the unmodified reference renderer in tests/test_rgb_render.py.
"""
import numpy as np
from PIL import Image, ImageDraw
from boxlcd_b200 import spec as _spec

ROBOT_COLORS = ((0.9, 0.4, 0.4), (0.5, 0.3, 0.5))
OBJECT_COLORS = ((0.5, 0.4, 0.9), (0.3, 0.3, 0.5))
F = np.float32


def body_shapes(sp, variant=0):
  """[(kind, radius, verts fp32 [n, 2], (color1, color2))] per dynamic body in draw order (= dynbodies order) for the shape
  variant bitmask `variant` (bit b picks body b's second shape: Object2/3's random circle-or-box, world_env.py:273-274)."""
  out = []
  for b in range(sp.n_bodies):
    bd = sp.bodies[b]
    sd = bd.shape[(int(variant) >> b) & 1 if bd.n_variants > 1 else 0]
    colors = OBJECT_COLORS if bd.role == _spec.ROLE_OBJECT else ROBOT_COLORS
    if sd.kind == _spec.SHAPE_CIRCLE:
      out.append(('circle', float(F(sd.radius)), None, colors))
    elif sd.kind == _spec.SHAPE_BOX:
      hx, hy = F(sd.verts[0][0]), F(sd.verts[0][1])
      out.append(('poly', 0.0, np.array([(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)], F), colors))
    else:
      ps = [(F(sd.verts[i][0]), F(sd.verts[i][1])) for i in range(sd.n_verts)]
      out.append(('poly', 0.0, np.array([ps[i] for i in _spec.hull_order(ps)], F), colors))
  return out


def _ink(c):
  return tuple(int(255.0 * (1 - x)) for x in c)


def render_rgb(shapes, pose, world_w, width, height):
  """One frame.  shapes: body_shapes(...); pose [B, 4] float32 (x, y, sin, cos).  -> uint8 [height, width, 3]"""
  image = Image.new('RGB', (width, height))
  draw = ImageDraw.Draw(image)
  draw.rectangle([0, 0, width, height], fill=(1, 1, 1))
  pose = np.asarray(pose, F)
  for (kind, radius, verts, (c1, c2)), (px, py, s, c) in zip(shapes, pose):
    if kind == 'circle':
      pos = np.array([px, py], np.float64)
      topleft = (pos - radius) / world_w * width
      botright = (pos + radius) / world_w * width
      draw.ellipse(topleft.tolist() + botright.tolist(), fill=_ink(c1), outline=_ink(c2))
    else:
      x = F(F(F(c * verts[:, 0]) - F(s * verts[:, 1])) + px)
      y = F(F(F(s * verts[:, 0]) + F(c * verts[:, 1])) + py)
      pts = np.stack([x, y], -1).astype(np.float64) / world_w
      pts = (width * pts).tolist()
      draw.polygon(tuple(tuple(xy) for xy in pts), fill=_ink(c1), outline=_ink(c2))
  image = image.transpose(method=Image.FLIP_TOP_BOTTOM)
  return 255 - np.asarray(image)


def human_frame(high_res, lcd):
  """The picture `render(mode='human')` hands to the viewer (world_env.py:525-531): [8x colour view | 1 px black | LCD
  frame blown up 8x].  lcd: bool [H, W] (mode '1') or uint8 [H, W, 3] (mode 'RGB')."""
  high_res = np.asarray(high_res).astype(np.uint8)
  lcd = np.asarray(lcd)
  if lcd.ndim == 3:
    low_res = lcd.astype(np.uint8).repeat(8, 0).repeat(8, 1)
  else:
    low_res = 255 * lcd.astype(np.uint8)[..., None].repeat(8, 0).repeat(8, 1).repeat(3, 2)
  return np.concatenate([high_res, np.zeros_like(low_res)[:, :1], low_res], axis=1)
