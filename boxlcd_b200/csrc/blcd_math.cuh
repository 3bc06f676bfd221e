// blcd_math.cuh -- fp32 2-D algebra for the batched world kernels (host+device inline).
// Operation order follows Box2D 2.3.x b2Math.h (the arithmetic behind the reference's b2World.Step,
// boxLCD/world_env.py:446-452) so that results stay within fp32 round-off of the CPU oracle.
#pragma once
#include "blcd_profile.h"
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define BLCD_HD __host__ __device__ __forceinline__
#define BLCD_HDN __host__ __device__ __noinline__
#else
#define BLCD_HD inline
#define BLCD_HDN
#endif

namespace BLCD_NS {

constexpr float kPi = 3.14159265359f;
constexpr float kEps = FLT_EPSILON;
constexpr float kMaxFloat = FLT_MAX;
constexpr float kLinearSlop = 0.005f;
constexpr float kAngularSlop = 2.0f / 180.0f * kPi;
constexpr float kPolygonRadius = 2.0f * kLinearSlop;
constexpr float kAabbExtension = 0.1f;
constexpr float kAabbMultiplier = 2.0f;
constexpr float kVelocityThreshold = 1.0f;
constexpr float kMaxLinearCorrection = 0.2f;
constexpr float kMaxAngularCorrection = 8.0f / 180.0f * kPi;
constexpr float kMaxTranslation = 2.0f;
constexpr float kMaxTranslationSq = kMaxTranslation * kMaxTranslation;
constexpr float kMaxRotation = 0.5f * kPi;
constexpr float kMaxRotationSq = kMaxRotation * kMaxRotation;
constexpr float kBaumgarte = 0.2f;
constexpr float kToiBaumgarte = 0.75f;
constexpr float kTimeToSleep = 0.5f;
constexpr float kLinSleepTol = 0.01f;
constexpr float kAngSleepTol = 2.0f / 180.0f * kPi;
constexpr int kMaxSubSteps = 8;

struct V2 {
  float x, y;
};
BLCD_HD V2 mk(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
BLCD_HD V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
BLCD_HD V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
BLCD_HD V2 operator-(V2 a) { return mk(-a.x, -a.y); }
BLCD_HD V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
BLCD_HD void operator+=(V2& a, V2 b) { a.x += b.x; a.y += b.y; }
BLCD_HD void operator-=(V2& a, V2 b) { a.x -= b.x; a.y -= b.y; }
BLCD_HD void operator*=(V2& a, float s) { a.x *= s; a.y *= s; }
BLCD_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
BLCD_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
BLCD_HD V2 cross(V2 a, float s) { return mk(s * a.y, -s * a.x); }
BLCD_HD V2 cross(float s, V2 a) { return mk(-s * a.y, s * a.x); }
BLCD_HD float len(V2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
BLCD_HD float len2(V2 a) { return a.x * a.x + a.y * a.y; }
BLCD_HD float dist2(V2 a, V2 b) { V2 c = a - b; return dot(c, c); }
BLCD_HD float fminb(float a, float b) { return a < b ? a : b; }   // b2Min
BLCD_HD float fmaxb(float a, float b) { return a > b ? a : b; }   // b2Max
BLCD_HD float clampb(float a, float lo, float hi) { return fmaxb(lo, fminb(a, hi)); }
BLCD_HD float absb(float a) { return a > 0.0f ? a : -a; }
// b2Vec2::Normalize
BLCD_HD float normalize(V2& v) {
  float l = len(v);
  if (l < kEps) return 0.0f;
  float inv = 1.0f / l;
  v.x *= inv; v.y *= inv;
  return l;
}

struct Rot {
  float s, c;
};
#ifdef __CUDA_ARCH__
// b2Rot::Set.  One out-of-line copy: sincosf expands to ~100 instructions (fast path + large-argument slow path) and
// is needed in ~30 places; inlining it everywhere made the kernel instruction-fetch bound.
__device__ __noinline__ Rot rot_of(float a) {
  Rot q;
  sincosf(a, &q.s, &q.c);
  return q;
}
#else
inline Rot rot_of(float a) {
  Rot q;
  q.s = sinf(a); q.c = cosf(a);
  return q;
}
#endif
BLCD_HD Rot rot_identity() { Rot q; q.s = 0.0f; q.c = 1.0f; return q; }
BLCD_HD V2 rmul(Rot q, V2 v) { return mk(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
BLCD_HD V2 rmulT(Rot q, V2 v) { return mk(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
BLCD_HD Rot rmulT(Rot q, Rot r) { Rot o; o.s = q.c * r.s - q.s * r.c; o.c = q.c * r.c + q.s * r.s; return o; }

struct Xf {
  V2 p;
  Rot q;
};
BLCD_HD Xf xf_identity() { Xf t; t.p = mk(0.0f, 0.0f); t.q = rot_identity(); return t; }
BLCD_HD V2 xmul(const Xf& T, V2 v) { return mk((T.q.c * v.x - T.q.s * v.y) + T.p.x, (T.q.s * v.x + T.q.c * v.y) + T.p.y); }
BLCD_HD V2 xmulT(const Xf& T, V2 v) {
  float px = v.x - T.p.x, py = v.y - T.p.y;
  return mk(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
BLCD_HD Xf xmulT(const Xf& A, const Xf& B) { Xf C; C.q = rmulT(A.q, B.q); C.p = rmulT(A.q, B.p - A.p); return C; }
// body transform from centre of mass c, angle a and local centre lc (b2Body::SynchronizeTransform)
BLCD_HD Xf xf_of(V2 c, float a, V2 lc) { Xf t; t.q = rot_of(a); t.p = c - rmul(t.q, lc); return t; }

struct Sweep {
  V2 lc, c0, c;
  float a0, a, alpha0;
};
BLCD_HD Xf sweep_xf(const Sweep& s, float beta) {
  Xf t;
  t.p = (1.0f - beta) * s.c0 + beta * s.c;
  float angle = (1.0f - beta) * s.a0 + beta * s.a;
  t.q = rot_of(angle);
  t.p -= rmul(t.q, s.lc);
  return t;
}
BLCD_HD void sweep_advance(Sweep& s, float alpha) {
  float beta = (alpha - s.alpha0) / (1.0f - s.alpha0);
  s.c0 += beta * (s.c - s.c0);
  s.a0 += beta * (s.a - s.a0);
  s.alpha0 = alpha;
}
BLCD_HD void sweep_normalize(Sweep& s) {
  float twoPi = 2.0f * kPi;
  float d = twoPi * floorf(s.a0 / twoPi);
  s.a0 -= d;
  s.a -= d;
}

struct Box {  // AABB
  V2 lo, hi;
};
BLCD_HD bool box_contains(const Box& a, const Box& b) { return a.lo.x <= b.lo.x && a.lo.y <= b.lo.y && b.hi.x <= a.hi.x && b.hi.y <= a.hi.y; }
BLCD_HD bool box_overlap(const Box& a, const Box& b) {
  V2 d1 = b.lo - a.hi, d2 = a.lo - b.hi;
  if (d1.x > 0.0f || d1.y > 0.0f) return false;
  if (d2.x > 0.0f || d2.y > 0.0f) return false;
  return true;
}

}  // namespace BLCD_NS
