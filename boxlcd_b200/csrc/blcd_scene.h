// blcd_scene.h -- device-side constant tables of one boxLCD scene and the layout of the per-world state in HBM.
//
// Built on the host from the `blcd_spec` POD (include/boxlcd_b200.h), which itself mirrors what the reference hands
// to pybox2d in WorldEnv.reset/_reset_bodies (boxLCD/world_env.py:197-316) and world_defs.py.  Shape hulls, normals,
// centroids and mass data are derived in fp32 the way b2PolygonShape::Set / SetAsBox / ComputeMass and
// b2Body::ResetMassData do (Box2D 2.3.x), because those values feed every later fp32 operation.
#pragma once
#include "blcd_math.cuh"
#include "boxlcd_b200.h"

namespace BLCD_NS {

enum { SH_CIRCLE = 0, SH_EDGE = 1, SH_POLY = 2 };
constexpr int kSlotWords = 16;   // one persistent manifold slot
constexpr int kBodyWords = 13;   // cx cy a vx vy w sleepTime fat(lo.x lo.y hi.x hi.y) px py (b2Body::m_xf.p)
constexpr int kJointWords = 7;   // impulse xyz, motorImpulse, motorSpeed, limitState, referenceAngle
// misc words: flags, inv_dt0, ep_t, rng draws [, shape variants].  With at most 8 bodies the flag word holds the awake
// bits (0..7), the shape-variant bits (8..15) and e_newFixture (16); the large profile keeps the variants in a word of
// their own.
constexpr bool kVariantInFlags = kMaxBodies <= 8;
constexpr int kMiscWords = kVariantInFlags ? 4 : 5;
constexpr uint32_t kBodyMask = (1u << kMaxBodies) - 1u;
constexpr uint32_t kAwakeMask = kBodyMask;                                   // bit b: dynamic body b awake
constexpr uint32_t kVariantShift = 8;                                        // (small profile) bits 8..15: shape variant of body b
constexpr uint32_t kNewFixtureBit = kVariantInFlags ? (1u << 16) : (1u << 31); // b2World::e_newFixture: run FindNewContacts at the next Step

struct DShape {
  int32_t type, count;
  float radius, rmax;   // rmax: largest distance of a core vertex from the body's centre of mass (TOI pre-filter)
  V2 v[BLCD_MAX_VERTS];
  V2 n[BLCD_MAX_VERTS];
  V2 centroid;
};

struct DBody {
  DShape shape[2];
  float invMass[2], invI[2];
  V2 lc[2];
  float friction, restitution, linDamp, angDamp;
  uint32_t cat, mask;
  int32_t nvar, role, parent, root, rand_angle;
  int32_t obs[4];
  int32_t njedge;
  int32_t jedge[kMaxJoints];  // joints attached to this body, newest first (b2Body::m_jointList order)
  double extent, joint_angle, anchor_a[2], anchor_b[2];
};

struct DJoint {
  int32_t a, b;  // dynamic body indices
  V2 la, lb;
  float lower, upper, maxTorque;
  int32_t enableLimit, enableMotor, act;
  double speed;
};

struct DPair {
  uint8_t fa, fb;  // fixture (= proxy) indices, fa < fb; walls are 0..nw-1, dynamic body b is nw+b
};

struct DScene {
  int32_t nb, nj, nw, np, has_robot;
  int32_t world_w, world_h, lcd_w, lcd_h, S, P, A;
  int32_t pobs[kMaxObs];
  int32_t nsub, vel_iters, pos_iters, ep_len, rules;
  uint32_t flags;
  float dt;
  V2 gravity;
  int32_t maxm;  // manifold slots per world
  int32_t align_mode;  // block-level phase alignment points: 0 none, 1 velocity loop, 2 + position loop, 3 + sub-step start and TOI, 4 + write-back, 5 + env step
  // per-world state layout in HBM: word i of world w lives at state[i * n_worlds + w]
  int32_t off_misc, off_joint, off_clist, clist_words, off_slots, off_cnt, state_words;
  // per-thread shared-memory layout (words), see blcd_world.cuh
  int32_t h_vel, h_pos, h_mass, h_joint, h_con, hot_words;
  // phase pipeline (blcd_pipeline.cuh): per-world scratch in HBM, word i of world w at scratch[i * n_worlds + w]:
  // body rows (v, c/a, inverse masses of the nb dynamic rows), joint records, solve bookkeeping (counts, joint order,
  // islands), pre-solve poses for the TOI sweeps, contact records
  int32_t x_rows, x_jr, x_misc, x_c0, x_cr, scratch_words;
  DShape wall[BLCD_MAX_WALLS];
  Box wallFat[BLCD_MAX_WALLS];
  V2 wall_n[BLCD_MAX_WALLS];      // unit normal of the wall's line and its offset (n . x = d), for the TOI pre-filter
  float wall_d[BLCD_MAX_WALLS];
  DBody body[kMaxBodies];
  DJoint joint[kMaxJoints];
  DPair pair[kMaxPairs];
};

// hot (shared-memory) record sizes
constexpr int kHotJoint = 18;        // rA rB K(6) motorMass impulse(3) motorImpulse motorSpeed packed referenceAngle
constexpr int kHotConHdr = 10;       // normal(2) friction packed K(3) normalMass(3)
constexpr int kHotConPt = 9;         // rA rB normalMass tangentMass bias ni ti
constexpr int kHotCon = kHotConHdr + 2 * kHotConPt;

// ---------------------------------------------------------------------------------------------------------------------
// host-side builder (fp32, no FMA contraction: compile with -ffp-contract=off)
namespace host {

inline float f32(double x) { return (float)x; }

inline void shape_circle(DShape& s, float r) {
  s = DShape();
  s.type = SH_CIRCLE; s.count = 1; s.radius = r;
}
inline void shape_edge(DShape& s, V2 a, V2 b) {
  s = DShape();
  s.type = SH_EDGE; s.count = 2; s.radius = kPolygonRadius; s.v[0] = a; s.v[1] = b;
}
inline void shape_box(DShape& s, float hx, float hy) {  // b2PolygonShape::SetAsBox
  s = DShape();
  s.type = SH_POLY; s.count = 4; s.radius = kPolygonRadius;
  s.v[0] = mk(-hx, -hy); s.v[1] = mk(hx, -hy); s.v[2] = mk(hx, hy); s.v[3] = mk(-hx, hy);
  s.n[0] = mk(0.0f, -1.0f); s.n[1] = mk(1.0f, 0.0f); s.n[2] = mk(0.0f, 1.0f); s.n[3] = mk(-1.0f, 0.0f);
}
inline void shape_polygon(DShape& s, const V2* in, int count) {  // b2PolygonShape::Set
  s = DShape();
  s.type = SH_POLY; s.radius = kPolygonRadius;
  V2 ps[BLCD_MAX_VERTS];
  int n = 0;
  for (int i = 0; i < count && i < BLCD_MAX_VERTS; ++i) {
    bool unique = true;
    for (int j = 0; j < n; ++j)
      if (dist2(in[i], ps[j]) < (0.5f * kLinearSlop) * (0.5f * kLinearSlop)) { unique = false; break; }
    if (unique) ps[n++] = in[i];
  }
  int i0 = 0;
  for (int i = 1; i < n; ++i)
    if (ps[i].x > ps[i0].x || (ps[i].x == ps[i0].x && ps[i].y < ps[i0].y)) i0 = i;
  int hull[BLCD_MAX_VERTS], m = 0, ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) { ie = j; continue; }
      V2 r = ps[ie] - ps[hull[m]], v = ps[j] - ps[hull[m]];
      float c = cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && len2(v) > len2(r)) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  s.count = m;
  for (int i = 0; i < m; ++i) s.v[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    V2 e = s.v[i + 1 < m ? i + 1 : 0] - s.v[i];
    s.n[i] = cross(e, 1.0f);
    normalize(s.n[i]);
  }
  V2 c = mk(0.0f, 0.0f);
  float area = 0.0f;
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < m; ++i) {  // ComputeCentroid, reference point at the origin
    V2 p2 = s.v[i], p3 = s.v[i + 1 < m ? i + 1 : 0];
    float D = cross(p2, p3);
    float ta = 0.5f * D;
    area += ta;
    c += (ta * inv3) * (mk(0.0f, 0.0f) + p2 + p3);
  }
  c *= 1.0f / area;
  s.centroid = c;
}

inline void mass_of(const DShape& s, float density, float* invMass, float* invI, V2* lc) {
  float mass = 0.0f, I = 0.0f;
  V2 center = mk(0.0f, 0.0f);
  if (s.type == SH_CIRCLE) {
    mass = density * kPi * s.radius * s.radius;
    I = mass * (0.5f * s.radius * s.radius + dot(center, center));
  } else {
    V2 ctr = mk(0.0f, 0.0f), ref = mk(0.0f, 0.0f);
    float area = 0.0f, acc = 0.0f;
    for (int i = 0; i < s.count; ++i) ref += s.v[i];
    ref *= 1.0f / s.count;
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < s.count; ++i) {
      V2 e1 = s.v[i] - ref, e2 = s.v[i + 1 < s.count ? i + 1 : 0] - ref;
      float D = cross(e1, e2);
      float ta = 0.5f * D;
      area += ta;
      ctr += (ta * k_inv3) * (e1 + e2);
      float intx2 = e1.x * e1.x + e2.x * e1.x + e2.x * e2.x;
      float inty2 = e1.y * e1.y + e2.y * e1.y + e2.y * e2.y;
      acc += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    mass = density * area;
    ctr *= 1.0f / area;
    center = ctr + ref;
    I = density * acc;
    I += mass * (dot(center, center) - dot(ctr, ctr));
  }
  // b2Body::ResetMassData for a single-fixture dynamic body
  V2 localCenter = mk(0.0f, 0.0f);
  float m = 0.0f, im = 0.0f, ii = 0.0f;
  if (density > 0.0f) {
    m = mass;
    localCenter = mass * center;
  }
  if (m > 0.0f) { im = 1.0f / m; localCenter *= im; } else { m = 1.0f; im = 1.0f; }
  if (density > 0.0f && I > 0.0f) {
    I -= m * dot(localCenter, localCenter);
    ii = 1.0f / I;
  }
  *invMass = im; *invI = ii; *lc = localCenter;
}

inline void make_shape(DShape& s, const blcd_shape_def& sd) {
  if (sd.kind == BLCD_SHAPE_CIRCLE) shape_circle(s, f32(sd.radius));
  else if (sd.kind == BLCD_SHAPE_BOX) shape_box(s, f32(sd.verts[0][0]), f32(sd.verts[0][1]));
  else {
    V2 vs[BLCD_MAX_VERTS];
    for (int i = 0; i < sd.n_verts; ++i) vs[i] = mk(f32(sd.verts[i][0]), f32(sd.verts[i][1]));
    shape_polygon(s, vs, sd.n_verts);
  }
}

// manifold slots per world: enough for every touching pair seen in long random rollouts of the reference scenes
// (tests/test_hostsim_vs_oracle.py measures it); overflow is counted in BLCD_CNT_OVERFLOW, never silent
inline int default_manifold_slots(int n_bodies) {
  int m = n_bodies <= 2 ? 4 : (n_bodies <= 4 ? 8 : (n_bodies == 5 ? 12 : (n_bodies <= 8 ? 16 : 32)));
  return m < kMaxSlots ? m : kMaxSlots;
}

// returns nullptr on success, else an error message
inline const char* build_scene(DScene& sc, const blcd_spec& sp, int maxm) {
  sc = DScene();
  if (sp.n_bodies < 1 || sp.n_bodies > kMaxBodies) return "n_bodies out of range for the " BLCD_PROFILE_NAME " profile";
  if (sp.n_joints < 0 || sp.n_joints > kMaxJoints) return "n_joints out of range for the " BLCD_PROFILE_NAME " profile";
  if (sp.n_walls < 0 || sp.n_walls > BLCD_MAX_WALLS) return "n_walls out of range";
  if (sp.lcd_w < 1 || sp.lcd_w > kRowBits || sp.lcd_h < 1 || sp.lcd_h > 64) return "frame size out of range for the " BLCD_PROFILE_NAME " profile";
  if (sp.obs_size != 4 * sp.n_bodies) return "obs_size must be 4 * n_bodies";
  sc.nb = sp.n_bodies; sc.nj = sp.n_joints; sc.nw = sp.n_walls; sc.has_robot = sp.has_robot;
  sc.world_w = sp.world_w; sc.world_h = sp.world_h; sc.lcd_w = sp.lcd_w; sc.lcd_h = sp.lcd_h;
  sc.S = sp.obs_size; sc.P = sp.pobs_size; sc.A = sp.act_size;
  for (int i = 0; i < kMaxObs; ++i) sc.pobs[i] = sp.pobs_index[i];   // kMaxObs <= BLCD_MAX_OBS
  sc.nsub = sp.n_substeps; sc.vel_iters = sp.vel_iters; sc.pos_iters = sp.pos_iters; sc.ep_len = sp.ep_len;
  sc.rules = sp.raster_rules; sc.flags = sp.flags; sc.dt = f32(sp.dt);
  sc.gravity = mk(f32(sp.gravity[0]), f32(sp.gravity[1]));
  for (int i = 0; i < sc.nw; ++i) {
    shape_edge(sc.wall[i], mk(f32(sp.walls[i][0]), f32(sp.walls[i][1])), mk(f32(sp.walls[i][2]), f32(sp.walls[i][3])));
    // static body at the origin: fat AABB = edge AABB +- polygonRadius +- aabbExtension (b2Fixture::CreateProxies)
    V2 a = sc.wall[i].v[0], b = sc.wall[i].v[1];
    V2 lo = mk(fminb(a.x, b.x), fminb(a.y, b.y)), hi = mk(fmaxb(a.x, b.x), fmaxb(a.y, b.y));
    V2 r = mk(kPolygonRadius, kPolygonRadius), e = mk(kAabbExtension, kAabbExtension);
    sc.wallFat[i].lo = (lo - r) - e;
    sc.wallFat[i].hi = (hi + r) + e;
    V2 ed = b - a;
    float el = len(ed);
    sc.wall_n[i] = el > 0.0f ? mk(-ed.y / el, ed.x / el) : mk(0.0f, 0.0f);
    sc.wall_d[i] = dot(sc.wall_n[i], a);
  }
  for (int b = 0; b < sc.nb; ++b) {
    const blcd_body_def& bd = sp.bodies[b];
    DBody& d = sc.body[b];
    d.nvar = bd.n_variants < 1 ? 1 : (bd.n_variants > 2 ? 2 : bd.n_variants);
    for (int k = 0; k < d.nvar; ++k) {
      make_shape(d.shape[k], bd.shape[k]);
      mass_of(d.shape[k], f32(bd.density), &d.invMass[k], &d.invI[k], &d.lc[k]);
      d.shape[k].rmax = 0.0f;
      for (int i = 0; i < d.shape[k].count; ++i) d.shape[k].rmax = fmaxb(d.shape[k].rmax, len(d.shape[k].v[i] - d.lc[k]));
    }
    if (d.nvar == 1) { d.shape[1] = d.shape[0]; d.invMass[1] = d.invMass[0]; d.invI[1] = d.invI[0]; d.lc[1] = d.lc[0]; }
    d.friction = f32(bd.friction); d.restitution = f32(bd.restitution);
    d.linDamp = f32(bd.linear_damping); d.angDamp = f32(bd.angular_damping);
    d.cat = bd.category_bits; d.mask = bd.mask_bits;
    d.role = bd.role; d.parent = bd.parent; d.root = bd.root; d.rand_angle = bd.rand_angle;
    for (int k = 0; k < 4; ++k) {
      d.obs[k] = bd.obs_index[k];
      if (d.obs[k] < 0 || d.obs[k] >= sc.S) return "obs_index out of range";
    }
    d.extent = bd.extent; d.joint_angle = bd.joint_angle;
    d.anchor_a[0] = bd.anchor_a[0]; d.anchor_a[1] = bd.anchor_a[1];
    d.anchor_b[0] = bd.anchor_b[0]; d.anchor_b[1] = bd.anchor_b[1];
    d.njedge = 0;
  }
  for (int j = 0; j < sc.nj; ++j) {
    const blcd_joint_def& jd = sp.joints[j];
    DJoint& d = sc.joint[j];
    if (jd.body_a < 0 || jd.body_a >= sc.nb || jd.body_b < 0 || jd.body_b >= sc.nb) return "joint body index out of range";
    d.a = jd.body_a; d.b = jd.body_b;
    d.la = mk(f32(jd.anchor_a[0]), f32(jd.anchor_a[1]));
    d.lb = mk(f32(jd.anchor_b[0]), f32(jd.anchor_b[1]));
    d.lower = f32(jd.lower); d.upper = f32(jd.upper); d.maxTorque = f32(jd.max_motor_torque);
    d.enableLimit = jd.enable_limit; d.enableMotor = jd.enable_motor; d.act = jd.act_index; d.speed = jd.speed;
    if (d.act >= sc.A) return "act_index out of range";
  }
  // joint edge lists, newest joint first
  for (int j = sc.nj - 1; j >= 0; --j) {
    DBody& a = sc.body[sc.joint[j].a];
    DBody& b = sc.body[sc.joint[j].b];
    a.jedge[a.njedge++] = j;
    b.jedge[b.njedge++] = j;
  }
  // candidate pairs in (proxy a, proxy b) ascending order: b2Body::ShouldCollide + b2ContactFilter::ShouldCollide
  int nf = sc.nw + sc.nb;
  sc.np = 0;
  for (int a = 0; a < nf; ++a) {
    for (int b = a + 1; b < nf; ++b) {
      if (b < sc.nw) continue;  // wall-wall: neither dynamic
      uint32_t catA = a < sc.nw ? 0x0001u : sc.body[a - sc.nw].cat, maskA = a < sc.nw ? 0xFFFFu : sc.body[a - sc.nw].mask;
      uint32_t catB = sc.body[b - sc.nw].cat, maskB = sc.body[b - sc.nw].mask;
      if (!((maskA & catB) != 0 && (catA & maskB) != 0)) continue;
      bool connected = false;
      if (a >= sc.nw)
        for (int j = 0; j < sc.nj; ++j) {
          int ja = sc.joint[j].a + sc.nw, jb = sc.joint[j].b + sc.nw;
          if ((ja == a && jb == b) || (ja == b && jb == a)) connected = true;
        }
      if (connected) continue;
      if (sc.np >= kMaxPairs) return "too many collidable fixture pairs for the " BLCD_PROFILE_NAME " profile";
      sc.pair[sc.np].fa = (uint8_t)a;
      sc.pair[sc.np].fb = (uint8_t)b;
      ++sc.np;
    }
  }
  sc.maxm = maxm;
  sc.align_mode = 4;
  sc.off_misc = kBodyWords * sc.nb;
  sc.off_joint = sc.off_misc + kMiscWords;
  sc.off_clist = sc.off_joint + kJointWords * sc.nj;
  sc.clist_words = 1 + (sc.np + 3) / 4;
  sc.off_slots = sc.off_clist + sc.clist_words;
  sc.off_cnt = sc.off_slots + kSlotWords * sc.maxm;
  sc.state_words = sc.off_cnt + BLCD_N_COUNTERS;
  // shared-memory layout: velocities and positions of nb dynamic rows + 1 static row, inverse-mass rows
  sc.h_vel = 0;
  sc.h_pos = sc.h_vel + 3 * (sc.nb + 1);
  sc.h_mass = sc.h_pos + 3 * (sc.nb + 1);
  sc.h_joint = 0;   // joint and contact records are thread-local (blcd_world.cuh), not in shared memory
  sc.h_con = 0;
  sc.hot_words = sc.h_mass + 2 * (sc.nb + 1);
  sc.x_rows = 0;
  sc.x_jr = sc.x_rows + 8 * sc.nb;
  sc.x_misc = sc.x_jr + kHotJoint * sc.nj;
  sc.x_c0 = sc.x_misc + 3 + 2 * ((sc.nj + 3) / 4) + (sc.nb + 3) / 4;   // misc ends with one word that survives from sub-step to sub-step: the number of position sweeps the world last needed
  sc.x_cr = sc.x_c0 + 3 * sc.nb;
  sc.scratch_words = sc.x_cr + kHotCon * sc.maxm;
  return nullptr;
}

}  // namespace host

}  // namespace BLCD_NS
