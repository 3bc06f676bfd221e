"""Stand-ins for the three pybox2d shape classes the reference's scene specs construct (`Box2D.b2.circleShape`,
`polygonShape`, `edgeShape`; world_defs.py:2-3, world_env.py:10).  They only carry the definition; hulls, normals and
mass data are derived inside libboxlcd_b200 in float32, the way Box2D does.  Values are rounded to float32 on
construction because that is what a b2Shape stores."""
import numpy as np

_f32 = lambda x: float(np.float32(x))


class circleShape:
  def __init__(self, radius=0.0, pos=(0.0, 0.0)):
    self.radius = _f32(radius)
    self.pos = (_f32(pos[0]), _f32(pos[1]))
    if self.pos != (0.0, 0.0):
      raise NotImplementedError('off-centre circle fixtures are not used by any boxLCD scene and are not supported')

  def __repr__(self):
    return f'circleShape(radius={self.radius})'


class polygonShape:
  """polygonShape(box=(hx, hy)) or polygonShape(vertices=[(x, y), ...]) (<= 8 vertices)."""
  def __init__(self, box=None, vertices=None):
    if (box is None) == (vertices is None):
      raise ValueError('give exactly one of box= / vertices=')
    if box is not None:
      self.box = (_f32(box[0]), _f32(box[1]))
      self.input_vertices = None
    else:
      self.box = None
      self.input_vertices = [(_f32(x), _f32(y)) for x, y in vertices]
      if not 3 <= len(self.input_vertices) <= 8:
        raise ValueError('polygonShape needs 3..8 vertices')

  @property
  def vertices(self):
    """stored vertex order: SetAsBox order for boxes, CCW gift-wrapped hull starting at the right-most point otherwise"""
    if self.box is not None:
      hx, hy = self.box
      return [(-hx, -hy), (hx, -hy), (hx, hy), (-hx, hy)]
    from boxlcd_b200.spec import hull_order
    return [self.input_vertices[i] for i in hull_order(self.input_vertices)]

  def __repr__(self):
    return f'polygonShape(box={self.box})' if self.box is not None else f'polygonShape(vertices={self.input_vertices})'


class edgeShape:
  def __init__(self, vertices):
    (x1, y1), (x2, y2) = vertices
    self.vertices = [(_f32(x1), _f32(y1)), (_f32(x2), _f32(y2))]
