/*
 * b2_oracle.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of the reference's hot path around `b2World.Step` for batches of independent worlds:
 *   WorldEnv.reset / _reset_bodies   boxLCD/world_env.py:197-385   -> reset_world()
 *   WorldEnv.step                    boxLCD/world_env.py:431-458   -> step_world()
 *   WorldEnv._get_obs                boxLCD/world_env.py:387-429   -> observe_world()
 *   examples/collect.py rollout loop examples/collect.py:31-39     -> blcd_oracle_worlds_rollout()
 * The Box2D arithmetic lives in b2_world.h / b2_collide.h / b2_math.h (restated from upstream Box2D 2.3.x, the
 * third-party C++ behind pybox2d `Box2D==2.3.10`, requirements.txt:17; sources not under /root/reference).
 *
 * Neither pybox2d nor Box2D sources exist in this image and the reference ships no tests for this path; the physics is
 * PINNED against the episodes the reference author recorded with real pybox2d (the gifs under assets/envs/): passive scenes frame
 * for frame, robot scenes (joints, motors, limits) for 9-12 s of random actions each (tests/test_gif_episodes.py,
 * oracle/README.md).  Also pinned: the LCD renderer (lcd_oracle.c vs the unmodified reference renderer) and the reset
 * distribution / observation layout (vs the reference run under stubs, tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call into this library.
 */
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include "b2_world.h"
#include "boxlcd_b200.h"

extern "C" {
typedef struct {
  int32_t kind;
  int32_t n;
  float radius;
  float verts[8][2];
} lcd_shape;
void blcd_oracle_lcd(const lcd_shape* shapes, int per_world_shapes, int n_bodies, const float* poses, int64_t n, int world_w,
                     int lcd_w, int lcd_h, int rules, uint32_t* bits);
}

namespace {

using namespace b2o;

// Philox4x32-10 (Salmon et al., SC'11), counter = (block, global world index lo, hi, 0), key = (seed lo, seed hi).
// Same generator and draw order as the CUDA path so that both sample identical resets and actions.
struct Philox {
  uint32_t key[2];
  uint32_t widx[2];
  uint32_t draws = 0;  // 32-bit values consumed so far
  uint32_t buf[4];
  uint32_t buf_block = 0xFFFFFFFFu;

  void block(uint32_t blk) {
    uint32_t c[4] = {blk, widx[0], widx[1], 0u};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
      uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
      uint32_t n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
      uint32_t n3 = (uint32_t)p0;
      c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    memcpy(buf, c, sizeof(buf));
    buf_block = blk;
  }
  uint32_t next() {
    uint32_t blk = draws >> 2;
    if (blk != buf_block) block(blk);
    return buf[draws++ & 3u];
  }
  double uniform01() { return (double)next() * (1.0 / 4294967296.0); }
  double uniform(double lo, double hi) { return lo + (hi - lo) * uniform01(); }
};

struct Env {
  World world;
  Philox rng;
  uint32_t variant = 0;  // bit b = shape variant of dynamic body b
  int32_t ep_t = 0;
};

struct Batch {
  blcd_spec spec;
  std::vector<Env> envs;
};

inline float f32(double x) { return (float)x; }

Shape make_shape(const blcd_shape_def& sd) {
  if (sd.kind == BLCD_SHAPE_CIRCLE) return MakeCircle(f32(sd.radius));
  if (sd.kind == BLCD_SHAPE_BOX) return MakeBox(f32(sd.verts[0][0]), f32(sd.verts[0][1]));
  Vec2 vs[kMaxPolygonVertices];
  for (int i = 0; i < sd.n_verts; ++i) vs[i] = Vec2(f32(sd.verts[i][0]), f32(sd.verts[i][1]));
  return MakePolygon(vs, sd.n_verts);
}

WorldFlags flags_of(const blcd_spec& sp) {
  WorldFlags f;
  f.damping_2_3_0 = (sp.flags & BLCD_FLAG_DAMPING_2_3_0) != 0;
  f.refface_2_3_0 = (sp.flags & BLCD_FLAG_REFFACE_2_3_0) != 0;
  f.no_toi = (sp.flags & BLCD_FLAG_NO_TOI) != 0;
  f.no_sleep = (sp.flags & BLCD_FLAG_NO_SLEEP) != 0;
  return f;
}

// world_env.py:190-195 (new b2World), :311-314 (walls), then bodies / joints in the reference's creation order.
// pose[b] = (x, y, angle) of dynamic body b at creation.
void build_world(Env& env, const blcd_spec& sp, const float (*pose)[3]) {
  env.world = World();
  World& w = env.world;
  w.flags = flags_of(sp);
  w.gravity = Vec2(f32(sp.gravity[0]), f32(sp.gravity[1]));
  for (int i = 0; i < sp.n_walls; ++i) {
    Shape e = MakeEdge(Vec2(f32(sp.walls[i][0]), f32(sp.walls[i][1])), Vec2(f32(sp.walls[i][2]), f32(sp.walls[i][3])));
    w.CreateBody(kStatic, Vec2(0.0f, 0.0f), 0.0f, e, 0.0f, 0.2f, 0.0f, 0x0001, 0xFFFF, 0.0f, 0.0f);
  }
  int nj = 0;
  for (int b = 0; b < sp.n_bodies; ++b) {
    const blcd_body_def& bd = sp.bodies[b];
    int var = (bd.n_variants > 1) ? (int)((env.variant >> b) & 1u) : 0;
    Shape s = make_shape(bd.shape[var]);
    w.CreateBody(kDynamic, Vec2(pose[b][0], pose[b][1]), pose[b][2], s, f32(bd.density), f32(bd.friction), f32(bd.restitution),
                 (uint16_t)bd.category_bits, (uint16_t)bd.mask_bits, f32(bd.linear_damping), f32(bd.angular_damping));
    if (bd.role == BLCD_ROLE_CHILD) {
      const blcd_joint_def& jd = sp.joints[nj++];
      RevoluteJoint j;
      j.bodyA = sp.n_walls + jd.body_a;
      j.bodyB = sp.n_walls + jd.body_b;
      j.localAnchorA = Vec2(f32(jd.anchor_a[0]), f32(jd.anchor_a[1]));
      j.localAnchorB = Vec2(f32(jd.anchor_b[0]), f32(jd.anchor_b[1]));
      // pybox2d's revoluteJointDef(bodyA=, bodyB=, localAnchorA=, ...) (world_env.py:255-266) fills referenceAngle with
      // bodyB.angle - bodyA.angle at definition time: the limits [lower, upper] are relative to the pose the robot is
      // assembled in.  Pinned by the recorded robot episodes (tests/test_gif_episodes.py); with referenceAngle = 0 the
      // legs of the urchin snap into a 2-rad fan within a few frames, which the recordings do not show.
      j.referenceAngle = pose[jd.body_b][2] - pose[jd.body_a][2];
      j.enableLimit = jd.enable_limit != 0;
      j.enableMotor = jd.enable_motor != 0;
      j.lowerAngle = f32(jd.lower);
      j.upperAngle = f32(jd.upper);
      j.maxMotorTorque = f32(jd.max_motor_torque);
      j.motorSpeed = 0.0f;
      w.CreateRevoluteJoint(j);
    }
  }
  env.ep_t = 0;
}

double mapto(double a, double lo, double hi) { return ((a + 1.0) / 2.0 * (hi - lo)) + lo; }
double rmapto(double a, double lo, double hi) { return ((a - lo) / (hi - lo) * 2.0) + -1.0; }

// world_env.py:197-304: sample the initial poses (float64 host arithmetic, float32 where pybox2d stores or adds b2Vec2s)
void sample_poses(Env& env, const blcd_spec& sp, float (*pose)[3]) {
  Philox& rng = env.rng;
  const double W = sp.world_w, H = sp.world_h;
  double angle64[BLCD_MAX_BODIES] = {0};
  env.variant = 0;
  for (int b = 0; b < sp.n_bodies; ++b) {
    const blcd_body_def& bd = sp.bodies[b];
    if (bd.role == BLCD_ROLE_ROOT) {
      double rangex = 1.0 - (2.0 * bd.extent / W), rangey = 1.0 - (2.0 * bd.extent / H);
      double x = mapto(rng.uniform(-rangex, rangex), 0.0, W);
      double y = mapto(rng.uniform(-rangey, -rangey), 0.0, H);
      double s = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
      double c = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
      double ang = atan2(s, c);
      if (!bd.rand_angle) ang = 0.0;
      angle64[b] = ang;
      pose[b][0] = f32(x); pose[b][1] = f32(y); pose[b][2] = f32(ang);
    } else if (bd.role == BLCD_ROLE_CHILD) {
      // world_env.py:230-252: children hang off their parent by the joint anchors; the joint angle is relative to the ROOT
      double mangle = angle64[bd.root] + bd.joint_angle;
      mangle = atan2(sin(mangle), cos(mangle));
      angle64[b] = mangle;
      double pangle = angle64[bd.parent];
      double aax = cos(pangle) * bd.anchor_a[0] - sin(pangle) * bd.anchor_a[1];
      double aay = sin(pangle) * bd.anchor_a[0] + cos(pangle) * bd.anchor_a[1];
      double abx = cos(mangle) * bd.anchor_b[0] - sin(mangle) * bd.anchor_b[1];
      double aby = sin(mangle) * bd.anchor_b[0] + cos(mangle) * bd.anchor_b[1];
      // b2Vec2 + sequence, b2Vec2 - sequence: float32 operands and results (pybox2d b2Vec2 operators)
      float px = pose[bd.parent][0] + f32(aax), py = pose[bd.parent][1] + f32(aay);
      px = px - f32(abx); py = py - f32(aby);
      pose[b][0] = px; pose[b][1] = py; pose[b][2] = f32(mangle);
    } else {
      if (bd.n_variants > 1) env.variant |= (rng.next() & 1u) << b;  // np.random.randint(2), world_env.py:274
      double rangex = 1.0 - (2.0 * bd.extent / W), rangey = 1.0 - (2.0 * bd.extent / H);
      double x = mapto(rng.uniform(-rangex, rangex), 0.0, W);
      double y = sp.has_robot ? mapto(rng.uniform(-rangey, -0.25), 0.0, H) : mapto(rng.uniform(-rangey, rangey), 0.0, H);
      double ang = 0.0;
      if (bd.rand_angle) {
        double s = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
        double c = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
        ang = atan2(s, c);
      }
      pose[b][0] = f32(x); pose[b][1] = f32(y); pose[b][2] = f32(ang);
    }
  }
}

// world_env.py:306-385.  full_state (normalized, may be null): applied with two SetTransform calls per body
// (`body.position = ...` then `body.angle = ...`), velocities stay zero.
void reset_world(Env& env, const blcd_spec& sp, const float* full_state) {
  float pose[BLCD_MAX_BODIES][3];
  sample_poses(env, sp, pose);
  build_world(env, sp, pose);
  if (full_state) {
    World& w = env.world;
    for (int pass = 0; pass < 2; ++pass) {  // objects first, then robot bodies (world_env.py:330-380)
      for (int b = 0; b < sp.n_bodies; ++b) {
        const blcd_body_def& bd = sp.bodies[b];
        if ((bd.role == BLCD_ROLE_OBJECT) != (pass == 0)) continue;
        double x = mapto((double)full_state[bd.obs_index[0]], 0.0, (double)sp.world_w);
        double y = mapto((double)full_state[bd.obs_index[1]], 0.0, (double)sp.world_h);
        double ang = atan2((double)full_state[bd.obs_index[3]], (double)full_state[bd.obs_index[2]]);
        int bi = sp.n_walls + b;
        w.SetTransform(bi, Vec2(f32(x), f32(y)), w.bodies[bi].sweep.a);
        w.SetTransform(bi, w.bodies[bi].xf.p, f32(ang));
      }
    }
  }
}

// world_env.py:431-452
void step_world(Env& env, const blcd_spec& sp, const float* action) {
  env.ep_t += 1;
  World& w = env.world;
  for (int j = 0; j < sp.n_joints; ++j) {
    const blcd_joint_def& jd = sp.joints[j];
    if (jd.act_index < 0) continue;
    double a = action ? (double)action[jd.act_index] : 0.0;
    a = a < -1.0 ? -1.0 : (a > 1.0 ? 1.0 : a);
    w.SetMotorSpeed(j, f32(jd.speed * a));
  }
  for (int s = 0; s < sp.n_substeps; ++s) w.Step(f32(sp.dt), sp.vel_iters, sp.pos_iters);
}

void draw_action(Env& env, const blcd_spec& sp, float* action) {
  for (int k = 0; k < sp.act_size; ++k) action[k] = f32(env.rng.uniform(-1.0, 1.0));
}

void get_bodies(const Env& env, const blcd_spec& sp, float* out) {
  for (int b = 0; b < sp.n_bodies; ++b) {
    const Body& bd = env.world.bodies[sp.n_walls + b];
    float* o = out + b * BLCD_BODY_STATE;
    o[0] = bd.xf.p.x; o[1] = bd.xf.p.y; o[2] = bd.sweep.a; o[3] = bd.v.x; o[4] = bd.v.y; o[5] = bd.w;
  }
}

void get_poses(const Env& env, const blcd_spec& sp, float* out) {
  for (int b = 0; b < sp.n_bodies; ++b) {
    const Body& bd = env.world.bodies[sp.n_walls + b];
    float* o = out + b * 4;
    o[0] = bd.xf.p.x; o[1] = bd.xf.p.y; o[2] = bd.xf.q.s; o[3] = bd.xf.q.c;
  }
}

void fill_lcd_shapes(const Env& env, const blcd_spec& sp, lcd_shape* out) {
  for (int b = 0; b < sp.n_bodies; ++b) {
    const Shape& s = env.world.bodies[sp.n_walls + b].shape;
    lcd_shape& o = out[b];
    memset(&o, 0, sizeof(o));
    if (s.type == kCircle) { o.kind = 0; o.radius = s.radius; }
    else {
      o.kind = 1; o.n = s.count;
      for (int i = 0; i < s.count; ++i) { o.verts[i][0] = s.v[i].x; o.verts[i][1] = s.v[i].y; }
    }
  }
}

// world_env.py:387-429 (float64 numpy arithmetic on float32 body state; stored here as float32 like the datasets)
void observe_world(const Env& env, const blcd_spec& sp, float* full_state, float* proprio, uint32_t* lcd_bits, uint8_t* done) {
  float fs[BLCD_MAX_OBS];
  for (int b = 0; b < sp.n_bodies; ++b) {
    const Body& bd = env.world.bodies[sp.n_walls + b];
    const int32_t* oi = sp.bodies[b].obs_index;
    fs[oi[0]] = f32(rmapto((double)bd.xf.p.x, 0.0, (double)sp.world_w));
    fs[oi[1]] = f32(rmapto((double)bd.xf.p.y, 0.0, (double)sp.world_h));
    fs[oi[2]] = f32(cos((double)bd.sweep.a));
    fs[oi[3]] = f32(sin((double)bd.sweep.a));
  }
  if (full_state) memcpy(full_state, fs, sizeof(float) * sp.obs_size);
  if (proprio) {
    if (sp.pobs_size == 0) proprio[0] = 0.0f;
    for (int i = 0; i < sp.pobs_size; ++i) proprio[i] = fs[sp.pobs_index[i]];
  }
  if (lcd_bits) {
    lcd_shape shapes[BLCD_MAX_BODIES];
    float poses[BLCD_MAX_BODIES * 4];
    fill_lcd_shapes(env, sp, shapes);
    get_poses(env, sp, poses);
    blcd_oracle_lcd(shapes, 0, sp.n_bodies, poses, 1, sp.world_w, sp.lcd_w, sp.lcd_h, sp.raster_rules, lcd_bits);
  }
  if (done) *done = env.ep_t >= sp.ep_len;
}

template <class F>
void parallel_for(int64_t n, int threads, F fn) {
  if (threads <= 1 || n <= 1) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
  if (threads > n) threads = (int)n;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([=]() {
      int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
      for (int64_t i = lo; i < hi; ++i) fn(i);
    });
  }
  for (auto& th : pool) th.join();
}

void write_counters(const Env& env, uint32_t* out) {
  const Counters& c = env.world.counters;
  out[BLCD_CNT_CONTACTS] = c.contacts; out[BLCD_CNT_POS_ITERS] = c.pos_iters; out[BLCD_CNT_TOI_EVENTS] = c.toi_events;
  out[BLCD_CNT_TOI_CALLS] = c.toi_calls; out[BLCD_CNT_SLEEP_STEPS] = c.sleep_steps; out[BLCD_CNT_OVERFLOW] = c.overflow;
  out[BLCD_CNT_MANIFOLD_POINTS] = c.manifold_points; out[BLCD_CNT_SUBSTEPS] = c.substeps;
}

}  // namespace

extern "C" {

void* blcd_oracle_worlds_new(const blcd_spec* spec, int64_t n, uint64_t seed, int64_t world_offset) {
  Batch* b = new Batch();
  b->spec = *spec;
  b->envs.resize((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    Env& e = b->envs[(size_t)i];
    uint64_t g = (uint64_t)(world_offset + i);
    e.rng.key[0] = (uint32_t)seed; e.rng.key[1] = (uint32_t)(seed >> 32);
    e.rng.widx[0] = (uint32_t)g; e.rng.widx[1] = (uint32_t)(g >> 32);
  }
  return b;
}

void blcd_oracle_worlds_free(void* h) { delete (Batch*)h; }

// new batch holding COPIES of the listed worlds (complete simulation state: contacts, warm-start impulses, RNG, ep_t)
void* blcd_oracle_worlds_select(void* h, const int64_t* idx, int64_t n) {
  Batch* src = (Batch*)h;
  Batch* b = new Batch();
  b->spec = src->spec;
  b->envs.reserve((size_t)n);
  for (int64_t i = 0; i < n; ++i) b->envs.push_back(src->envs[(size_t)idx[i]]);
  return b;
}

int blcd_oracle_worlds_reset(void* h, const int64_t* idx, int64_t n, const float* full_state, int threads) {
  Batch* b = (Batch*)h;
  if (!idx) n = (int64_t)b->envs.size();
  parallel_for(n, threads, [&](int64_t i) {
    int64_t w = idx ? idx[i] : i;
    reset_world(b->envs[(size_t)w], b->spec, full_state ? full_state + i * b->spec.obs_size : nullptr);
  });
  return 0;
}

// fresh worlds with bodies created at the given poses / velocities (the single-step parity protocol)
int blcd_oracle_worlds_set_bodies(void* h, const float* bodies, const uint32_t* variants, int threads) {
  Batch* b = (Batch*)h;
  const blcd_spec& sp = b->spec;
  parallel_for((int64_t)b->envs.size(), threads, [&](int64_t i) {
    Env& e = b->envs[(size_t)i];
    e.variant = variants ? variants[i] : 0u;
    float pose[BLCD_MAX_BODIES][3];
    const float* src = bodies + i * sp.n_bodies * BLCD_BODY_STATE;
    for (int k = 0; k < sp.n_bodies; ++k) { pose[k][0] = src[k * 6 + 0]; pose[k][1] = src[k * 6 + 1]; pose[k][2] = src[k * 6 + 2]; }
    build_world(e, sp, pose);
    for (int k = 0; k < sp.n_bodies; ++k) {
      Body& bd = e.world.bodies[sp.n_walls + k];
      bd.v = Vec2(src[k * 6 + 3], src[k * 6 + 4]);
      bd.w = src[k * 6 + 5];
    }
  });
  return 0;
}

int blcd_oracle_worlds_get_bodies(void* h, float* bodies) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) get_bodies(b->envs[i], b->spec, bodies + i * b->spec.n_bodies * BLCD_BODY_STATE);
  return 0;
}

int blcd_oracle_worlds_get_poses(void* h, float* poses, uint32_t* variants) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) {
    get_poses(b->envs[i], b->spec, poses + i * b->spec.n_bodies * 4);
    if (variants) variants[i] = b->envs[i].variant;
  }
  return 0;
}

int blcd_oracle_worlds_step(void* h, const float* actions, float* actions_out, int threads) {
  Batch* b = (Batch*)h;
  const blcd_spec& sp = b->spec;
  parallel_for((int64_t)b->envs.size(), threads, [&](int64_t i) {
    Env& e = b->envs[(size_t)i];
    float act[BLCD_MAX_OBS];
    if (actions) memcpy(act, actions + i * sp.act_size, sizeof(float) * sp.act_size);
    else draw_action(e, sp, act);
    if (actions_out) memcpy(actions_out + i * sp.act_size, act, sizeof(float) * sp.act_size);
    step_world(e, sp, act);
  });
  return 0;
}

int blcd_oracle_worlds_observe(void* h, float* full_state, float* proprio, uint32_t* lcd_bits, uint8_t* done, int threads) {
  Batch* b = (Batch*)h;
  const blcd_spec& sp = b->spec;
  int P = sp.pobs_size > 0 ? sp.pobs_size : 1;
  parallel_for((int64_t)b->envs.size(), threads, [&](int64_t i) {
    observe_world(b->envs[(size_t)i], sp, full_state ? full_state + i * sp.obs_size : nullptr, proprio ? proprio + i * P : nullptr,
                  lcd_bits ? lcd_bits + i * sp.lcd_h * BLCD_LCD_WORDS(sp.lcd_w) : nullptr, done ? done + i : nullptr);
  });
  return 0;
}

// examples/collect.py:31-39 for every world: outputs [n, T, ...]
int blcd_oracle_worlds_rollout(void* h, int32_t T, float* full_state, uint32_t* lcd_bits, float* actions, int threads) {
  Batch* b = (Batch*)h;
  const blcd_spec& sp = b->spec;
  parallel_for((int64_t)b->envs.size(), threads, [&](int64_t i) {
    Env& e = b->envs[(size_t)i];
    float act[BLCD_MAX_OBS];
    for (int t = 0; t < T; ++t) {
      int64_t o = i * T + t;
      observe_world(e, sp, full_state ? full_state + o * sp.obs_size : nullptr, nullptr, lcd_bits ? lcd_bits + o * sp.lcd_h * BLCD_LCD_WORDS(sp.lcd_w) : nullptr, nullptr);
      draw_action(e, sp, act);
      if (actions) memcpy(actions + o * sp.act_size, act, sizeof(float) * sp.act_size);
      step_world(e, sp, act);
    }
  });
  return 0;
}

int blcd_oracle_worlds_counters(void* h, uint32_t* counters) {
  Batch* b = (Batch*)h;
  for (size_t i = 0; i < b->envs.size(); ++i) write_counters(b->envs[i], counters + i * BLCD_N_COUNTERS);
  return 0;
}

// child placement algebra only (world_env.py:230-252): given the pose (x, y, angle) of every ROOT / OBJECT body in
// pose_in [n_bodies][3] (children ignored), writes all poses with the children placed.  angle is taken as float64 from
// the float32 input, like `root_angle` held in Python.
int blcd_oracle_place_children(const blcd_spec* sp, const double* pose_in, float* pose_out) {
  double angle64[BLCD_MAX_BODIES] = {0};
  float pose[BLCD_MAX_BODIES][3];
  for (int b = 0; b < sp->n_bodies; ++b) {
    const blcd_body_def& bd = sp->bodies[b];
    if (bd.role != BLCD_ROLE_CHILD) {
      angle64[b] = pose_in[b * 3 + 2];
      pose[b][0] = f32(pose_in[b * 3]); pose[b][1] = f32(pose_in[b * 3 + 1]); pose[b][2] = f32(pose_in[b * 3 + 2]);
    } else {
      double mangle = angle64[bd.root] + bd.joint_angle;
      mangle = atan2(sin(mangle), cos(mangle));
      angle64[b] = mangle;
      double pangle = angle64[bd.parent];
      double aax = cos(pangle) * bd.anchor_a[0] - sin(pangle) * bd.anchor_a[1];
      double aay = sin(pangle) * bd.anchor_a[0] + cos(pangle) * bd.anchor_a[1];
      double abx = cos(mangle) * bd.anchor_b[0] - sin(mangle) * bd.anchor_b[1];
      double aby = sin(mangle) * bd.anchor_b[0] + cos(mangle) * bd.anchor_b[1];
      float px = pose[bd.parent][0] + f32(aax), py = pose[bd.parent][1] + f32(aay);
      px = px - f32(abx); py = py - f32(aby);
      pose[b][0] = px; pose[b][1] = py; pose[b][2] = f32(mangle);
    }
  }
  for (int b = 0; b < sp->n_bodies; ++b) { pose_out[b * 3] = pose[b][0]; pose_out[b * 3 + 1] = pose[b][1]; pose_out[b * 3 + 2] = pose[b][2]; }
  return 0;
}

// shape table of world i in lcd_oracle layout (for render-only tests)
int blcd_oracle_worlds_lcd_shapes(void* h, int64_t i, void* out) {
  Batch* b = (Batch*)h;
  fill_lcd_shapes(b->envs[(size_t)i], b->spec, (lcd_shape*)out);
  return 0;
}

// mass properties of dynamic body k of world i: mass, inertia about the centre of mass, local centre x, y
int blcd_oracle_worlds_mass(void* h, int64_t i, int k, float* out) {
  Batch* b = (Batch*)h;
  const Body& bd = b->envs[(size_t)i].world.bodies[b->spec.n_walls + k];
  out[0] = bd.mass; out[1] = bd.I; out[2] = bd.sweep.localCenter.x; out[3] = bd.sweep.localCenter.y;
  return 0;
}

}  // extern "C"
