"""WorldDef + config -> `blcd_spec` (include/boxlcd_b200.h): the flat, pointer-free description of one boxLCD scene
that libboxlcd_b200 turns into device tables.

Follows the reference's construction order so indices mean the same thing on both sides:
  * dynamic bodies in `dynbodies` insertion order (world_env.py:197-304): per robot root, then children in
    `robot.joints` order, then objects -- this is also the LCD draw order (world_env.py:478);
  * joints in creation order (world_env.py:255-267);
  * walls in creation order (world_env.py:311-314), or the single floor edge of walls=0 (:316);
  * observation keys sorted alphabetically (world_env.py:119-126), action keys likewise (:138-141).
"""
import ctypes as C
import numpy as np
from boxlcd_b200 import utils
from boxlcd_b200.shapes import circleShape, polygonShape

MAX_BODIES, MAX_JOINTS, MAX_WALLS, MAX_VERTS = 18, 17, 4, 8
MAX_OBS = 4 * MAX_BODIES
SHAPE_CIRCLE, SHAPE_BOX, SHAPE_POLYGON = 0, 1, 2
ROLE_OBJECT, ROLE_ROOT, ROLE_CHILD = 0, 1, 2
RASTER_PIL12, RASTER_PIL9 = 0, 1
FLAG_DAMPING_2_3_0, FLAG_REFFACE_2_3_0, FLAG_NO_TOI, FLAG_NO_SLEEP = 1, 2, 4, 8


class ShapeDef(C.Structure):
  _fields_ = [('kind', C.c_int32), ('n_verts', C.c_int32), ('radius', C.c_double), ('verts', (C.c_double * 2) * MAX_VERTS)]


class BodyDef(C.Structure):
  _fields_ = [('n_variants', C.c_int32), ('role', C.c_int32), ('shape', ShapeDef * 2),
              ('density', C.c_double), ('friction', C.c_double), ('restitution', C.c_double),
              ('linear_damping', C.c_double), ('angular_damping', C.c_double),
              ('category_bits', C.c_uint32), ('mask_bits', C.c_uint32),
              ('rand_angle', C.c_int32), ('parent', C.c_int32), ('root', C.c_int32),
              ('extent', C.c_double), ('joint_angle', C.c_double),
              ('anchor_a', C.c_double * 2), ('anchor_b', C.c_double * 2), ('obs_index', C.c_int32 * 4)]


class JointDef(C.Structure):
  _fields_ = [('body_a', C.c_int32), ('body_b', C.c_int32), ('enable_limit', C.c_int32), ('enable_motor', C.c_int32),
              ('anchor_a', C.c_double * 2), ('anchor_b', C.c_double * 2), ('lower', C.c_double), ('upper', C.c_double),
              ('max_motor_torque', C.c_double), ('speed', C.c_double), ('act_index', C.c_int32), ('_pad', C.c_int32)]


class Spec(C.Structure):
  _fields_ = [('n_bodies', C.c_int32), ('n_joints', C.c_int32), ('n_walls', C.c_int32), ('has_robot', C.c_int32),
              ('bodies', BodyDef * MAX_BODIES), ('joints', JointDef * MAX_JOINTS), ('walls', (C.c_double * 4) * MAX_WALLS),
              ('gravity', C.c_double * 2), ('world_w', C.c_int32), ('world_h', C.c_int32), ('lcd_w', C.c_int32), ('lcd_h', C.c_int32),
              ('obs_size', C.c_int32), ('pobs_size', C.c_int32), ('act_size', C.c_int32), ('pobs_index', C.c_int32 * MAX_OBS),
              ('n_substeps', C.c_int32), ('vel_iters', C.c_int32), ('pos_iters', C.c_int32), ('dt', C.c_double),
              ('ep_len', C.c_int32), ('raster_rules', C.c_int32), ('flags', C.c_uint32), ('_pad', C.c_int32)]


def hull_order(ps):
  """Vertex order b2PolygonShape::Set stores for input points `ps` (already float32-rounded): gift wrapping, CCW,
  starting at the right-most point (lowest y on ties).  Arithmetic in float32 like Box2D."""
  f = np.float32
  n = len(ps)
  i0 = 0
  for i in range(1, n):
    if ps[i][0] > ps[i0][0] or (ps[i][0] == ps[i0][0] and ps[i][1] < ps[i0][1]):
      i0 = i
  order, ih = [], i0
  while True:
    order.append(ih)
    ie = 0
    for j in range(1, n):
      if ie == ih:
        ie = j
        continue
      rx, ry = f(ps[ie][0]) - f(ps[ih][0]), f(ps[ie][1]) - f(ps[ih][1])
      vx, vy = f(ps[j][0]) - f(ps[ih][0]), f(ps[j][1]) - f(ps[ih][1])
      c = f(rx * vy) - f(ry * vx)
      if c < 0 or (c == 0 and f(vx * vx) + f(vy * vy) > f(rx * rx) + f(ry * ry)):
        ie = j
    ih = ie
    if ie == i0:
      break
  return order


def _fill_shape(sd, shape):
  if isinstance(shape, circleShape):
    sd.kind, sd.radius = SHAPE_CIRCLE, shape.radius
  elif isinstance(shape, polygonShape):
    if shape.box is not None:
      sd.kind = SHAPE_BOX
      sd.verts[0][0], sd.verts[0][1] = shape.box
    else:
      sd.kind, sd.n_verts = SHAPE_POLYGON, len(shape.input_vertices)
      for i, (x, y) in enumerate(shape.input_vertices):
        sd.verts[i][0], sd.verts[i][1] = x, y
  else:
    raise TypeError(f'unsupported shape {shape!r}')


class SceneLayout:
  """Everything the host wrapper needs to know about a compiled scene besides the C struct itself."""
  def __init__(self):
    self.spec = Spec()
    self.body_names = []      # dynbodies order
    self.joint_names = []
    self.obs_info = {}
    self.act_info = {}


def compile_spec(world_def, G, width, height):
  """world_def: WorldDef with robots already filled by ROBOT_FILLER.  G: AttrDict config.  width/height: WIDTH/HEIGHT."""
  if G.all_corners or G.compact_obs or G.root_offset or G.angular_offset:
    raise NotImplementedError('only the default observation layout (x, y, cos, sin per body) is built; '
                              'all_corners is unusable in the reference too (ipdb.set_trace at world_env.py:178)')
  if not G.use_speed:
    raise NotImplementedError('torque control is broken in the reference (act key :force vs lookup :torque, world_env.py:114,443) and is not built')
  if not G.walls and not world_def.robots:
    raise IndexError('walls=0 needs a robot to follow (world_env.py:382 indexes world_def.robots[0])')
  lay = SceneLayout()
  sp = lay.spec
  A = utils.A
  obs_info, act_info = {}, {}
  for obj in world_def.objects:
    obs_info[f'{obj.name}:x:p'] = A[0, width]
    obs_info[f'{obj.name}:y:p'] = A[0, height]
    obs_info[f'{obj.name}:cos'] = A[-1, 1]
    obs_info[f'{obj.name}:sin'] = A[-1, 1]
  for robot in world_def.robots:
    for part in ['root'] + list(robot.joints):
      obs_info[f'{robot.name}:{part}:x:p'] = A[0, width]
      obs_info[f'{robot.name}:{part}:y:p'] = A[0, height]
      obs_info[f'{robot.name}:{part}:cos'] = A[-1, 1]
      obs_info[f'{robot.name}:{part}:sin'] = A[-1, 1]
    for jn, joint in robot.joints.items():
      if joint.limits[0] != joint.limits[1]:
        act_info[f'{robot.name}:{jn}:speed'] = A[-1, 1]
  if not world_def.robots:
    act_info['dummy'] = A[-1, 1]
  lay.obs_info = obs_info = utils.sortdict(obs_info)
  lay.act_info = act_info = utils.sortdict(act_info)
  obs_keys, act_keys = list(obs_info), list(act_info)
  pobs_keys = utils.nfiltlist(obs_keys, 'object')

  def obs_idx(prefix):
    return [obs_keys.index(f'{prefix}:{s}') for s in ('x:p', 'y:p', 'cos', 'sin')]

  nb = nj = 0
  for robot in world_def.robots:
    if nb >= MAX_BODIES:
      raise NotImplementedError(f'scene needs more than {MAX_BODIES} dynamic bodies')
    root_idx = nb
    index_of = {'root': root_idx}
    bd = sp.bodies[nb]
    rb = robot.root_body
    bd.n_variants, bd.role = 1, ROLE_ROOT
    _fill_shape(bd.shape[0], rb.shape)
    bd.density = 1.0 if rb.density is None else rb.density
    bd.friction, bd.restitution = 1.0, 0.0            # world_env.py:203 hard-codes friction=1.0 for the root
    bd.linear_damping, bd.angular_damping = robot.linearDamping, robot.angularDamping
    bd.category_bits, bd.mask_bits = rb.categoryBits, rb.maskBits
    bd.rand_angle, bd.parent, bd.root, bd.extent = int(robot.rand_angle), -1, root_idx, robot.bound
    bd.obs_index[:] = obs_idx(f'{robot.name}:root')
    lay.body_names.append(f'{robot.name}:root')
    nb += 1
    for jn, joint in robot.joints.items():
      if nb >= MAX_BODIES or nj >= MAX_JOINTS:
        raise NotImplementedError(f'scene needs more than {MAX_BODIES} dynamic bodies / {MAX_JOINTS} joints')
      body = robot.bodies[jn]
      bd = sp.bodies[nb]
      bd.n_variants, bd.role = 1, ROLE_CHILD
      _fill_shape(bd.shape[0], body.shape)
      bd.density, bd.friction, bd.restitution = 1.0, body.friction, 0.0   # world_env.py:238: density=1, restitution=0
      bd.category_bits, bd.mask_bits = body.categoryBits, body.maskBits
      bd.parent, bd.root = index_of[joint.parent], root_idx
      bd.joint_angle = joint.angle
      bd.anchor_a[:] = [float(joint.anchorA[0]), float(joint.anchorA[1])]
      bd.anchor_b[:] = [float(joint.anchorB[0]), float(joint.anchorB[1])]
      bd.obs_index[:] = obs_idx(f'{robot.name}:{jn}')
      index_of[jn] = nb
      lay.body_names.append(f'{robot.name}:{jn}')
      jd = sp.joints[nj]
      jd.body_a, jd.body_b = bd.parent, nb
      jd.enable_limit, jd.enable_motor = int(bool(joint.limited)), 1
      jd.anchor_a[:] = bd.anchor_a[:]
      jd.anchor_b[:] = bd.anchor_b[:]
      jd.lower, jd.upper = joint.limits
      jd.max_motor_torque, jd.speed = joint.torque, joint.speed
      key = f'{robot.name}:{jn}:speed'
      jd.act_index = act_keys.index(key) if key in act_info else -1
      lay.joint_names.append(f'{robot.name}:{jn}')
      nb += 1
      nj += 1
  for obj in world_def.objects:
    if nb >= MAX_BODIES:
      raise NotImplementedError(f'scene needs more than {MAX_BODIES} dynamic bodies')
    if obj.rangex is not None or obj.rangey is not None:
      raise NotImplementedError('Object.rangex/rangey overrides leave the sampling range undefined in the reference (world_env.py:278-281)')
    bd = sp.bodies[nb]
    bd.role = ROLE_OBJECT
    alts = {'circle': circleShape(radius=obj.size), 'box': polygonShape(box=(obj.size, obj.size))}
    if obj.shape == 'random':
      bd.n_variants = 2
      _fill_shape(bd.shape[0], alts['circle'])   # dict order of obj_shapes at world_env.py:273
      _fill_shape(bd.shape[1], alts['box'])
    else:
      bd.n_variants = 1
      _fill_shape(bd.shape[0], alts[obj.shape])
    bd.density, bd.friction, bd.restitution = obj.density, obj.friction, obj.restitution
    bd.linear_damping, bd.angular_damping = obj.linearDamping, obj.angularDamping
    bd.category_bits, bd.mask_bits = obj.categoryBits, 0xFFFF
    bd.rand_angle, bd.parent, bd.root, bd.extent = int(obj.rand_angle), -1, -1, obj.size
    bd.obs_index[:] = obs_idx(obj.name)
    lay.body_names.append(obj.name)
    nb += 1
  sp.n_bodies, sp.n_joints, sp.has_robot = nb, nj, int(len(world_def.robots) > 0)
  if G.walls:
    walls = [(0, 0, width, 0), (0, 0, 0, height), (width, 0, width, height), (0, height, width, height)]
  else:   # open world: one long floor edge, nothing else (world_env.py:316); frames keep showing [0, WIDTH) (:462 "TODO")
    walls = [(-1000 * width, 0, 1000 * width, 0)]
  sp.n_walls = len(walls)
  for i, w in enumerate(walls):
    sp.walls[i][:] = [float(x) for x in w]
  sp.gravity[:] = [float(world_def.gravity[0]), float(world_def.gravity[1])]
  sp.world_w, sp.world_h = int(width), int(height)
  sp.lcd_w, sp.lcd_h = int(G.lcd_base * G.wh_ratio), int(G.lcd_base)
  if sp.lcd_w > 64:
    raise NotImplementedError('frames wider than 64 px')
  sp.obs_size, sp.pobs_size, sp.act_size = len(obs_keys), len(pobs_keys), len(act_keys)
  for i, k in enumerate(pobs_keys):
    sp.pobs_index[i] = obs_keys.index(k)
  fps = G.fps
  sp.n_substeps, sp.dt = (3, 1.0 / (fps * 3)) if fps < 30 else (1, 1.0 / fps)
  sp.vel_iters, sp.pos_iters = 6 * 30, 2 * 30
  sp.ep_len = int(G.ep_len)
  sp.raster_rules = {'pil12': RASTER_PIL12, 'pil9': RASTER_PIL9}[G.get('raster_rules', 'pil12')]
  # Box2D revision switches.  pybox2d 2.3.10 behaves like Box2D 2.3.0 for damping (v *= clamp(1 - h d, 0, 1)) and like 2.3.1+
  # for the polygon reference-face rule: the recorded UrchinCube episode (cube with linearDamping 1.0) is reproduced for
  # 133 of 150 frames with FLAG_DAMPING_2_3_0 alone, 58 with the Pade form, 133 but with more stray pixels with both
  # 2.3.0 flags (tests/test_gif_episodes.py).
  sp.flags = int(G.get('b2_flags', FLAG_DAMPING_2_3_0))
  return lay
