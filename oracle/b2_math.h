/*
 * b2_math.h -- TEST INFRASTRUCTURE (part of the CPU oracle), not product code.
 * fp32 vector / rotation / sweep algebra restated from Box2D 2.3.x `Box2D/Common/b2Math.h` (third-party dependency of the
 * reference: requirements.txt:17 `Box2D==2.3.10`, not vendored under /root/reference).  Every operator rounds each
 * product and sum separately; the oracle is compiled with -ffp-contract=off like an x86-64 Box2D build.
 */
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace b2o {

constexpr float kPi = 3.14159265359f;
constexpr float kEpsilon = FLT_EPSILON;
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
constexpr float kMaxTranslationSquared = kMaxTranslation * kMaxTranslation;
constexpr float kMaxRotation = 0.5f * kPi;
constexpr float kMaxRotationSquared = kMaxRotation * kMaxRotation;
constexpr float kBaumgarte = 0.2f;
constexpr float kToiBaumgarte = 0.75f;
constexpr float kTimeToSleep = 0.5f;
constexpr float kLinearSleepTolerance = 0.01f;
constexpr float kAngularSleepTolerance = 2.0f / 180.0f * kPi;
constexpr int kMaxSubSteps = 8;
constexpr int kMaxTOIContacts = 32;
constexpr int kMaxPolygonVertices = 8;

struct Vec2 {
  float x, y;
  Vec2() : x(0.0f), y(0.0f) {}
  Vec2(float x_, float y_) : x(x_), y(y_) {}
  Vec2 operator-() const { return Vec2(-x, -y); }
  void operator+=(const Vec2& v) { x += v.x; y += v.y; }
  void operator-=(const Vec2& v) { x -= v.x; y -= v.y; }
  void operator*=(float a) { x *= a; y *= a; }
  float Length() const { return sqrtf(x * x + y * y); }
  float LengthSquared() const { return x * x + y * y; }
  float Normalize() {
    float length = Length();
    if (length < kEpsilon) return 0.0f;
    float inv = 1.0f / length;
    x *= inv; y *= inv;
    return length;
  }
};
inline Vec2 operator+(const Vec2& a, const Vec2& b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(const Vec2& a, const Vec2& b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator*(float s, const Vec2& a) { return Vec2(s * a.x, s * a.y); }
inline bool operator==(const Vec2& a, const Vec2& b) { return a.x == b.x && a.y == b.y; }
inline float Dot(const Vec2& a, const Vec2& b) { return a.x * b.x + a.y * b.y; }
inline float Cross(const Vec2& a, const Vec2& b) { return a.x * b.y - a.y * b.x; }
inline Vec2 Cross(const Vec2& a, float s) { return Vec2(s * a.y, -s * a.x); }
inline Vec2 Cross(float s, const Vec2& a) { return Vec2(-s * a.y, s * a.x); }
inline float DistanceSquared(const Vec2& a, const Vec2& b) { Vec2 c = a - b; return Dot(c, c); }
inline float Distance(const Vec2& a, const Vec2& b) { Vec2 c = a - b; return c.Length(); }
inline float Min(float a, float b) { return a < b ? a : b; }
inline float Max(float a, float b) { return a > b ? a : b; }
inline Vec2 Min(const Vec2& a, const Vec2& b) { return Vec2(Min(a.x, b.x), Min(a.y, b.y)); }
inline Vec2 Max(const Vec2& a, const Vec2& b) { return Vec2(Max(a.x, b.x), Max(a.y, b.y)); }
inline float Clamp(float a, float lo, float hi) { return Max(lo, Min(a, hi)); }
inline float Abs(float a) { return a > 0.0f ? a : -a; }

struct Vec3 {
  float x, y, z;
  Vec3() : x(0.0f), y(0.0f), z(0.0f) {}
  Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  Vec3 operator-() const { return Vec3(-x, -y, -z); }
  void operator+=(const Vec3& v) { x += v.x; y += v.y; z += v.z; }
  void operator*=(float s) { x *= s; y *= s; z *= s; }
};
inline float Dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 Cross(const Vec3& a, const Vec3& b) { return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct Mat22 {
  Vec2 ex, ey;
  Mat22 GetInverse() const {
    float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
    Mat22 B;
    float det = a * d - b * c;
    if (det != 0.0f) det = 1.0f / det;
    B.ex.x = det * d; B.ey.x = -det * b;
    B.ex.y = -det * c; B.ey.y = det * a;
    return B;
  }
  Vec2 Solve(const Vec2& b) const {
    float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    return Vec2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
  }
};
inline Vec2 Mul(const Mat22& A, const Vec2& v) { return Vec2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y); }

struct Mat33 {
  Vec3 ex, ey, ez;
  Vec3 Solve33(const Vec3& b) const {
    float det = Dot(ex, Cross(ey, ez));
    if (det != 0.0f) det = 1.0f / det;
    Vec3 x;
    x.x = det * Dot(b, Cross(ey, ez));
    x.y = det * Dot(ex, Cross(b, ez));
    x.z = det * Dot(ex, Cross(ey, b));
    return x;
  }
  Vec2 Solve22(const Vec2& b) const {
    float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    return Vec2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
  }
};

struct Rot {
  float s, c;
  Rot() : s(0.0f), c(1.0f) {}
  explicit Rot(float angle) { Set(angle); }
  void Set(float angle) { s = sinf(angle); c = cosf(angle); }
};
inline Vec2 Mul(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline Vec2 MulT(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
inline Rot MulT(const Rot& q, const Rot& r) {
  Rot qr;
  qr.s = q.c * r.s - q.s * r.c;
  qr.c = q.c * r.c + q.s * r.s;
  return qr;
}

struct Transform {
  Vec2 p;
  Rot q;
};
inline Vec2 Mul(const Transform& T, const Vec2& v) {
  float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
  float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
  return Vec2(x, y);
}
inline Vec2 MulT(const Transform& T, const Vec2& v) {
  float px = v.x - T.p.x, py = v.y - T.p.y;
  return Vec2(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
inline Transform MulT(const Transform& A, const Transform& B) {
  Transform C;
  C.q = MulT(A.q, B.q);
  C.p = MulT(A.q, B.p - A.p);
  return C;
}

struct Sweep {
  Vec2 localCenter, c0, c;
  float a0 = 0.0f, a = 0.0f, alpha0 = 0.0f;
  void GetTransform(Transform* xf, float beta) const {
    xf->p = (1.0f - beta) * c0 + beta * c;
    float angle = (1.0f - beta) * a0 + beta * a;
    xf->q.Set(angle);
    xf->p -= Mul(xf->q, localCenter);
  }
  void Advance(float alpha) {
    float beta = (alpha - alpha0) / (1.0f - alpha0);
    c0 += beta * (c - c0);
    a0 += beta * (a - a0);
    alpha0 = alpha;
  }
  void Normalize() {
    float twoPi = 2.0f * kPi;
    float d = twoPi * floorf(a0 / twoPi);
    a0 -= d;
    a -= d;
  }
};

struct AABB {
  Vec2 lower, upper;
  void Combine(const AABB& a, const AABB& b) { lower = Min(a.lower, b.lower); upper = Max(a.upper, b.upper); }
  bool Contains(const AABB& o) const {
    return lower.x <= o.lower.x && lower.y <= o.lower.y && o.upper.x <= upper.x && o.upper.y <= upper.y;
  }
};
inline bool TestOverlap(const AABB& a, const AABB& b) {
  Vec2 d1 = b.lower - a.upper, d2 = a.lower - b.upper;
  if (d1.x > 0.0f || d1.y > 0.0f) return false;
  if (d2.x > 0.0f || d2.y > 0.0f) return false;
  return true;
}

}  // namespace b2o
