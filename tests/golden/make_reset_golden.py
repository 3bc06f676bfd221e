"""Generate tests/golden/reset_golden.json from the UNMODIFIED reference (`/root/reference/boxLCD`) run under the
recording stubs of ref_harness.py: WorldEnv.__init__ metadata (key order, sizes, spaces), the fixture / joint definitions
the reference hands to pybox2d in reset(), and the initial poses of 12 seeded resets per env.

Run in the build container only:   python tests/golden/make_reset_golden.py
Pins: boxlcd_b200.spec.compile_spec (Appendix A constants), the child placement algebra and the reset ranges
(Appendix B) of both the oracle and the CUDA path (tests/test_reset_golden.py)."""
import json
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

ENVS = ['Dropbox', 'Bounce', 'Bounce2', 'Object2', 'Object3', 'Urchin', 'Luxo', 'UrchinCube', 'LuxoCube', 'UrchinBall', 'LuxoBall',
        'UrchinBalls', 'LuxoBalls', 'UrchinCubes', 'LuxoCubes', 'Crab', 'CrabCube', 'SpiderCube']


def shape_desc(shape):
  if isinstance(shape, ref_harness.circleShape):
    return {'kind': 'circle', 'radius': shape.radius}
  return {'kind': 'polygon', 'vertices': [list(map(float, v)) for v in shape.vertices]}


def main():
  boxLCD = ref_harness.ref_envs()
  out = {}
  for name in ENVS:
    env = boxLCD.env_map[name]()
    rec = {'obs_keys': env.obs_keys, 'act_keys': env.act_keys, 'pobs_idxs': [int(i) for i in env.pobs_idxs], 'obs_size': env.obs_size,
           'act_size': env.act_size, 'pobs_size': env.pobs_size, 'WIDTH': env.WIDTH, 'HEIGHT': env.HEIGHT,
           'lcd_shape': list(env.observation_space.spaces['lcd'].shape), 'ENV_DG': {k: (v if not isinstance(v, np.generic) else v.item()) for k, v in env.G.items()},
           'resets': []}
    np.random.seed(0)
    for seed in range(12):
      env.seed(seed)
      obs = env.reset()
      bodies, joints = [], []
      for bname, body in env.dynbodies.items():
        fx = body.fixtures[0]
        bodies.append({'name': bname, 'position': [float(body.position[0]), float(body.position[1])], 'angle': float(body.angle),
                       'angle64': float(body.kw.get('_angle64', body.angle)),
                       'density': float(fx.density), 'friction': float(getattr(fx, 'friction', 0.2)), 'restitution': float(getattr(fx, 'restitution', 0.0)),
                       'categoryBits': int(getattr(fx, 'categoryBits', 1)), 'maskBits': int(getattr(fx, 'maskBits', 0xFFFF)),
                       'linearDamping': float(body.kw.get('linearDamping', 0.0)), 'angularDamping': float(body.kw.get('angularDamping', 0.0)),
                       'shape': shape_desc(fx.shape)})
      names = list(env.dynbodies)
      bodyidx = {id(b): i for i, b in enumerate(env.dynbodies.values())}
      for jname, j in env.joints.items():
        jd = j.jd
        joints.append({'name': jname, 'bodyA': bodyidx[id(jd.bodyA)], 'bodyB': bodyidx[id(jd.bodyB)], 'localAnchorA': [float(x) for x in jd.localAnchorA],
                       'localAnchorB': [float(x) for x in jd.localAnchorB], 'enableMotor': bool(jd.enableMotor), 'enableLimit': bool(jd.enableLimit),
                       'maxMotorTorque': float(jd.maxMotorTorque), 'motorSpeed': float(jd.motorSpeed), 'lowerAngle': float(jd.lowerAngle), 'upperAngle': float(jd.upperAngle)})
      item = {'seed': seed, 'bodies': bodies, 'full_state': [float(x) for x in obs['full_state']]}
      if seed == 0:
        item['joints'] = joints
      else:
        for b in item['bodies']:   # constants are recorded once; later resets keep poses (and shapes, which may be random)
          for k in ('density', 'friction', 'restitution', 'categoryBits', 'maskBits', 'linearDamping', 'angularDamping'):
            b.pop(k)
      rec['resets'].append(item)
    out[name] = rec
    print(name, len(rec['resets']))
  json.dump(out, open(os.path.join(HERE, 'reset_golden.json'), 'w'))


if __name__ == '__main__':
  main()
