/*
 * b2_collide.h -- TEST INFRASTRUCTURE (part of the CPU oracle), not product code.
 * Shapes, mass data, narrow phase, GJK distance and conservative-advancement time of impact, restated from Box2D 2.3.x
 * (`Box2D/Collision/Shapes/b2{Circle,Edge,Polygon}Shape.cpp`, `b2CollideCircle.cpp`, `b2CollidePolygon.cpp`,
 * `b2CollideEdge.cpp`, `b2Collision.cpp`, `b2Distance.cpp`, `b2TimeOfImpact.cpp`) -- the C++ that pybox2d 2.3.10 wraps
 * and the reference calls through `b2World.Step` (boxLCD/world_env.py:446-452).  Box2D is not vendored in
 * /root/reference and pybox2d is not installable here: parity is pinned through the reference's recorded episodes instead (tests/test_gif_hires.py; see oracle/README.md).
 */
#pragma once
#include "b2_math.h"

namespace b2o {

enum ShapeType { kCircle = 0, kEdge = 1, kPolygon = 2 };

struct Shape {
  int type = kCircle;
  float radius = 0.0f;  // circle radius, or b2_polygonRadius skin for edges / polygons
  int count = 0;        // circle: 1 (v[0] = centre), edge: 2, polygon: n
  Vec2 v[kMaxPolygonVertices];
  Vec2 n[kMaxPolygonVertices];
  Vec2 centroid;
};

struct MassData {
  float mass = 0.0f, I = 0.0f;
  Vec2 center;
};

inline Shape MakeCircle(float r) {
  Shape s;
  s.type = kCircle; s.radius = r; s.count = 1; s.v[0] = Vec2(0.0f, 0.0f);
  return s;
}

inline Shape MakeEdge(Vec2 v1, Vec2 v2) {
  Shape s;
  s.type = kEdge; s.radius = kPolygonRadius; s.count = 2; s.v[0] = v1; s.v[1] = v2;
  return s;
}

// b2PolygonShape::SetAsBox(hx, hy)
inline Shape MakeBox(float hx, float hy) {
  Shape s;
  s.type = kPolygon; s.radius = kPolygonRadius; s.count = 4;
  s.v[0] = Vec2(-hx, -hy); s.v[1] = Vec2(hx, -hy); s.v[2] = Vec2(hx, hy); s.v[3] = Vec2(-hx, hy);
  s.n[0] = Vec2(0.0f, -1.0f); s.n[1] = Vec2(1.0f, 0.0f); s.n[2] = Vec2(0.0f, 1.0f); s.n[3] = Vec2(-1.0f, 0.0f);
  s.centroid = Vec2(0.0f, 0.0f);
  return s;
}

// b2PolygonShape.cpp: ComputeCentroid
inline Vec2 ComputeCentroid(const Vec2* vs, int count) {
  Vec2 c(0.0f, 0.0f);
  float area = 0.0f;
  Vec2 pRef(0.0f, 0.0f);
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < count; ++i) {
    Vec2 p1 = pRef, p2 = vs[i], p3 = i + 1 < count ? vs[i + 1] : vs[0];
    Vec2 e1 = p2 - p1, e2 = p3 - p1;
    float D = Cross(e1, e2);
    float triangleArea = 0.5f * D;
    area += triangleArea;
    c += triangleArea * inv3 * (p1 + p2 + p3);
  }
  c *= 1.0f / area;
  return c;
}

// b2PolygonShape::Set -- weld, gift-wrap hull, normals, centroid
inline Shape MakePolygon(const Vec2* vertices, int count) {
  Shape s;
  s.type = kPolygon; s.radius = kPolygonRadius;
  int n = count < kMaxPolygonVertices ? count : kMaxPolygonVertices;
  Vec2 ps[kMaxPolygonVertices];
  int tempCount = 0;
  for (int i = 0; i < n; ++i) {
    Vec2 v = vertices[i];
    bool unique = true;
    for (int j = 0; j < tempCount; ++j) {
      if (DistanceSquared(v, ps[j]) < ((0.5f * kLinearSlop) * (0.5f * kLinearSlop))) { unique = false; break; }
    }
    if (unique) ps[tempCount++] = v;
  }
  n = tempCount;
  int i0 = 0;
  float x0 = ps[0].x;
  for (int i = 1; i < n; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
  }
  int hull[kMaxPolygonVertices];
  int m = 0, ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < n; ++j) {
      if (ie == ih) { ie = j; continue; }
      Vec2 r = ps[ie] - ps[hull[m]];
      Vec2 v = ps[j] - ps[hull[m]];
      float c = Cross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && v.LengthSquared() > r.LengthSquared()) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  s.count = m;
  for (int i = 0; i < m; ++i) s.v[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    int i2 = i + 1 < m ? i + 1 : 0;
    Vec2 edge = s.v[i2] - s.v[i];
    s.n[i] = Cross(edge, 1.0f);
    s.n[i].Normalize();
  }
  s.centroid = ComputeCentroid(s.v, m);
  return s;
}

inline MassData ComputeMass(const Shape& s, float density) {
  MassData md;
  if (s.type == kCircle) {
    md.mass = density * kPi * s.radius * s.radius;
    md.center = s.v[0];
    md.I = md.mass * (0.5f * s.radius * s.radius + Dot(s.v[0], s.v[0]));
  } else if (s.type == kPolygon) {
    Vec2 center(0.0f, 0.0f);
    float area = 0.0f, I = 0.0f;
    Vec2 ref(0.0f, 0.0f);
    for (int i = 0; i < s.count; ++i) ref += s.v[i];
    ref *= 1.0f / s.count;
    const float k_inv3 = 1.0f / 3.0f;
    for (int i = 0; i < s.count; ++i) {
      Vec2 e1 = s.v[i] - ref;
      Vec2 e2 = i + 1 < s.count ? s.v[i + 1] - ref : s.v[0] - ref;
      float D = Cross(e1, e2);
      float triangleArea = 0.5f * D;
      area += triangleArea;
      center += triangleArea * k_inv3 * (e1 + e2);
      float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
      float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
      float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
      I += (0.25f * k_inv3 * D) * (intx2 + inty2);
    }
    md.mass = density * area;
    center *= 1.0f / area;
    md.center = center + ref;
    md.I = density * I;
    md.I += md.mass * (Dot(md.center, md.center) - Dot(center, center));
  }
  return md;
}

inline AABB ComputeAABB(const Shape& s, const Transform& xf) {
  AABB bb;
  if (s.type == kCircle) {
    Vec2 p = xf.p + Mul(xf.q, s.v[0]);
    bb.lower = Vec2(p.x - s.radius, p.y - s.radius);
    bb.upper = Vec2(p.x + s.radius, p.y + s.radius);
    return bb;
  }
  Vec2 lower = Mul(xf, s.v[0]), upper = lower;
  for (int i = 1; i < s.count; ++i) {
    Vec2 v = Mul(xf, s.v[i]);
    lower = Min(lower, v);
    upper = Max(upper, v);
  }
  Vec2 r(s.radius, s.radius);
  bb.lower = lower - r;
  bb.upper = upper + r;
  return bb;
}

// ---------------------------------------------------------------------------------------------------------------------
// manifolds
enum ManifoldType { kCircles = 0, kFaceA = 1, kFaceB = 2 };
enum FeatureType { kVertex = 0, kFace = 1 };

struct ContactID {
  uint8_t indexA = 0, indexB = 0, typeA = 0, typeB = 0;
  uint32_t key() const { return (uint32_t)indexA | ((uint32_t)indexB << 8) | ((uint32_t)typeA << 16) | ((uint32_t)typeB << 24); }
};

struct ManifoldPoint {
  Vec2 localPoint;
  float normalImpulse = 0.0f, tangentImpulse = 0.0f;
  ContactID id;
};

struct Manifold {
  ManifoldPoint points[2];
  Vec2 localNormal, localPoint;
  int type = kCircles;
  int pointCount = 0;
};

struct ClipVertex {
  Vec2 v;
  ContactID id;
};

inline void CollideCircles(Manifold* m, const Shape& A, const Transform& xfA, const Shape& B, const Transform& xfB) {
  m->pointCount = 0;
  Vec2 pA = Mul(xfA, A.v[0]), pB = Mul(xfB, B.v[0]);
  Vec2 d = pB - pA;
  float distSqr = Dot(d, d);
  float radius = A.radius + B.radius;
  if (distSqr > radius * radius) return;
  m->type = kCircles;
  m->localPoint = A.v[0];
  m->localNormal = Vec2(0.0f, 0.0f);
  m->pointCount = 1;
  m->points[0].localPoint = B.v[0];
  m->points[0].id = ContactID();
}

inline void CollidePolygonAndCircle(Manifold* m, const Shape& A, const Transform& xfA, const Shape& B, const Transform& xfB) {
  m->pointCount = 0;
  Vec2 c = Mul(xfB, B.v[0]);
  Vec2 cLocal = MulT(xfA, c);
  int normalIndex = 0;
  float separation = -kMaxFloat;
  float radius = A.radius + B.radius;
  for (int i = 0; i < A.count; ++i) {
    float s = Dot(A.n[i], cLocal - A.v[i]);
    if (s > radius) return;
    if (s > separation) { separation = s; normalIndex = i; }
  }
  int vertIndex1 = normalIndex;
  int vertIndex2 = vertIndex1 + 1 < A.count ? vertIndex1 + 1 : 0;
  Vec2 v1 = A.v[vertIndex1], v2 = A.v[vertIndex2];
  if (separation < kEpsilon) {
    m->pointCount = 1; m->type = kFaceA;
    m->localNormal = A.n[normalIndex];
    m->localPoint = 0.5f * (v1 + v2);
    m->points[0].localPoint = B.v[0];
    m->points[0].id = ContactID();
    return;
  }
  float u1 = Dot(cLocal - v1, v2 - v1);
  float u2 = Dot(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (DistanceSquared(cLocal, v1) > radius * radius) return;
    m->pointCount = 1; m->type = kFaceA;
    m->localNormal = cLocal - v1;
    m->localNormal.Normalize();
    m->localPoint = v1;
  } else if (u2 <= 0.0f) {
    if (DistanceSquared(cLocal, v2) > radius * radius) return;
    m->pointCount = 1; m->type = kFaceA;
    m->localNormal = cLocal - v2;
    m->localNormal.Normalize();
    m->localPoint = v2;
  } else {
    Vec2 faceCenter = 0.5f * (v1 + v2);
    float sep = Dot(cLocal - faceCenter, A.n[vertIndex1]);
    if (sep > radius) return;
    m->pointCount = 1; m->type = kFaceA;
    m->localNormal = A.n[vertIndex1];
    m->localPoint = faceCenter;
  }
  m->points[0].localPoint = B.v[0];
  m->points[0].id = ContactID();
}

inline int ClipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], const Vec2& normal, float offset, int vertexIndexA) {
  int numOut = 0;
  float distance0 = Dot(normal, vIn[0].v) - offset;
  float distance1 = Dot(normal, vIn[1].v) - offset;
  if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
  if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
  if (distance0 * distance1 < 0.0f) {
    float interp = distance0 / (distance0 - distance1);
    vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
    vOut[numOut].id.indexA = (uint8_t)vertexIndexA;
    vOut[numOut].id.indexB = vIn[0].id.indexB;
    vOut[numOut].id.typeA = kVertex;
    vOut[numOut].id.typeB = kFace;
    ++numOut;
  }
  return numOut;
}

inline float FindMaxSeparation(int* edgeIndex, const Shape& poly1, const Transform& xf1, const Shape& poly2, const Transform& xf2) {
  Transform xf = MulT(xf2, xf1);
  int bestIndex = 0;
  float maxSeparation = -kMaxFloat;
  for (int i = 0; i < poly1.count; ++i) {
    Vec2 n = Mul(xf.q, poly1.n[i]);
    Vec2 v1 = Mul(xf, poly1.v[i]);
    float si = kMaxFloat;
    for (int j = 0; j < poly2.count; ++j) {
      float sij = Dot(n, poly2.v[j] - v1);
      if (sij < si) si = sij;
    }
    if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
  }
  *edgeIndex = bestIndex;
  return maxSeparation;
}

// Box2D 2.3.0's b2EdgeSeparation / hill-climbing b2FindMaxSeparation (replaced by the brute-force loop above in 2.3.1).
// The reference's recorded Object2-cubes episode is reproduced further with the 2.3.0 rules (tests/test_gif_episodes.py),
// so this is what pybox2d 2.3.10 appears to vendor.
inline float EdgeSeparation230(const Shape& poly1, const Transform& xf1, int edge1, const Shape& poly2, const Transform& xf2) {
  Vec2 normal1World = Mul(xf1.q, poly1.n[edge1]);
  Vec2 normal1 = MulT(xf2.q, normal1World);
  int index = 0;
  float minDot = kMaxFloat;
  for (int i = 0; i < poly2.count; ++i) {
    float dot = Dot(poly2.v[i], normal1);
    if (dot < minDot) { minDot = dot; index = i; }
  }
  Vec2 v1 = Mul(xf1, poly1.v[edge1]);
  Vec2 v2 = Mul(xf2, poly2.v[index]);
  return Dot(v2 - v1, normal1World);
}

inline float FindMaxSeparation230(int* edgeIndex, const Shape& poly1, const Transform& xf1, const Shape& poly2, const Transform& xf2) {
  int count1 = poly1.count;
  Vec2 d = Mul(xf2, poly2.centroid) - Mul(xf1, poly1.centroid);
  Vec2 dLocal1 = MulT(xf1.q, d);
  int edge = 0;
  float maxDot = -kMaxFloat;
  for (int i = 0; i < count1; ++i) {
    float dot = Dot(poly1.n[i], dLocal1);
    if (dot > maxDot) { maxDot = dot; edge = i; }
  }
  float s = EdgeSeparation230(poly1, xf1, edge, poly2, xf2);
  int prevEdge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
  float sPrev = EdgeSeparation230(poly1, xf1, prevEdge, poly2, xf2);
  int nextEdge = edge + 1 < count1 ? edge + 1 : 0;
  float sNext = EdgeSeparation230(poly1, xf1, nextEdge, poly2, xf2);
  int bestEdge, increment;
  float bestSeparation;
  if (sPrev > s && sPrev > sNext) { increment = -1; bestEdge = prevEdge; bestSeparation = sPrev; }
  else if (sNext > s) { increment = 1; bestEdge = nextEdge; bestSeparation = sNext; }
  else { *edgeIndex = edge; return s; }
  for (;;) {
    if (increment == -1) edge = bestEdge - 1 >= 0 ? bestEdge - 1 : count1 - 1;
    else edge = bestEdge + 1 < count1 ? bestEdge + 1 : 0;
    s = EdgeSeparation230(poly1, xf1, edge, poly2, xf2);
    if (s > bestSeparation) { bestEdge = edge; bestSeparation = s; }
    else break;
  }
  *edgeIndex = bestEdge;
  return bestSeparation;
}

inline void FindIncidentEdge(ClipVertex c[2], const Shape& poly1, const Transform& xf1, int edge1, const Shape& poly2, const Transform& xf2) {
  Vec2 normal1 = MulT(xf2.q, Mul(xf1.q, poly1.n[edge1]));
  int index = 0;
  float minDot = kMaxFloat;
  for (int i = 0; i < poly2.count; ++i) {
    float dot = Dot(normal1, poly2.n[i]);
    if (dot < minDot) { minDot = dot; index = i; }
  }
  int i1 = index, i2 = i1 + 1 < poly2.count ? i1 + 1 : 0;
  c[0].v = Mul(xf2, poly2.v[i1]);
  c[0].id.indexA = (uint8_t)edge1; c[0].id.indexB = (uint8_t)i1; c[0].id.typeA = kFace; c[0].id.typeB = kVertex;
  c[1].v = Mul(xf2, poly2.v[i2]);
  c[1].id.indexA = (uint8_t)edge1; c[1].id.indexB = (uint8_t)i2; c[1].id.typeA = kFace; c[1].id.typeB = kVertex;
}

// refface_2_3_0: use the 2.3.0 reference-face hysteresis (0.98*sepA + 0.001) instead of sepA + 0.1*linearSlop
inline void CollidePolygons(Manifold* m, const Shape& polyA, const Transform& xfA, const Shape& polyB, const Transform& xfB, bool refface_2_3_0) {
  m->pointCount = 0;
  float totalRadius = polyA.radius + polyB.radius;
  int edgeA = 0;
  float separationA = refface_2_3_0 ? FindMaxSeparation230(&edgeA, polyA, xfA, polyB, xfB) : FindMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
  if (separationA > totalRadius) return;
  int edgeB = 0;
  float separationB = refface_2_3_0 ? FindMaxSeparation230(&edgeB, polyB, xfB, polyA, xfA) : FindMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
  if (separationB > totalRadius) return;
  const Shape *poly1, *poly2;
  Transform xf1, xf2;
  int edge1;
  bool flip;
  bool useB;
  if (refface_2_3_0) {
    useB = separationB > 0.98f * separationA + 0.001f;
  } else {
    const float k_tol = 0.1f * kLinearSlop;
    useB = separationB > separationA + k_tol;
  }
  if (useB) {
    poly1 = &polyB; poly2 = &polyA; xf1 = xfB; xf2 = xfA; edge1 = edgeB; m->type = kFaceB; flip = true;
  } else {
    poly1 = &polyA; poly2 = &polyB; xf1 = xfA; xf2 = xfB; edge1 = edgeA; m->type = kFaceA; flip = false;
  }
  ClipVertex incidentEdge[2];
  FindIncidentEdge(incidentEdge, *poly1, xf1, edge1, *poly2, xf2);
  int iv1 = edge1, iv2 = edge1 + 1 < poly1->count ? edge1 + 1 : 0;
  Vec2 v11 = poly1->v[iv1], v12 = poly1->v[iv2];
  Vec2 localTangent = v12 - v11;
  localTangent.Normalize();
  Vec2 localNormal = Cross(localTangent, 1.0f);
  Vec2 planePoint = 0.5f * (v11 + v12);
  Vec2 tangent = Mul(xf1.q, localTangent);
  Vec2 normal = Cross(tangent, 1.0f);
  v11 = Mul(xf1, v11);
  v12 = Mul(xf1, v12);
  float frontOffset = Dot(normal, v11);
  float sideOffset1 = -Dot(tangent, v11) + totalRadius;
  float sideOffset2 = Dot(tangent, v12) + totalRadius;
  ClipVertex clipPoints1[2], clipPoints2[2];
  int np = ClipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = ClipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  m->localNormal = localNormal;
  m->localPoint = planePoint;
  int pointCount = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = Dot(normal, clipPoints2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      ManifoldPoint* cp = m->points + pointCount;
      cp->localPoint = MulT(xf2, clipPoints2[i].v);
      cp->id = clipPoints2[i].id;
      if (flip) {
        ContactID cf = cp->id;
        cp->id.indexA = cf.indexB; cp->id.indexB = cf.indexA; cp->id.typeA = cf.typeB; cp->id.typeB = cf.typeA;
      }
      ++pointCount;
    }
  }
  m->pointCount = pointCount;
}

// b2CollideEdgeAndCircle for an edge without ghost vertices (the reference's walls, world_env.py:311-316)
inline void CollideEdgeAndCircle(Manifold* m, const Shape& edgeA, const Transform& xfA, const Shape& circleB, const Transform& xfB) {
  m->pointCount = 0;
  Vec2 Q = MulT(xfA, Mul(xfB, circleB.v[0]));
  Vec2 A = edgeA.v[0], B = edgeA.v[1];
  Vec2 e = B - A;
  float u = Dot(e, B - Q);
  float v = Dot(e, Q - A);
  float radius = edgeA.radius + circleB.radius;
  ContactID cf;
  cf.indexB = 0; cf.typeB = kVertex;
  if (v <= 0.0f) {
    Vec2 P = A;
    Vec2 d = Q - P;
    float dd = Dot(d, d);
    if (dd > radius * radius) return;
    cf.indexA = 0; cf.typeA = kVertex;
    m->pointCount = 1; m->type = kCircles;
    m->localNormal = Vec2(0.0f, 0.0f);
    m->localPoint = P;
    m->points[0].id = cf;
    m->points[0].localPoint = circleB.v[0];
    return;
  }
  if (u <= 0.0f) {
    Vec2 P = B;
    Vec2 d = Q - P;
    float dd = Dot(d, d);
    if (dd > radius * radius) return;
    cf.indexA = 1; cf.typeA = kVertex;
    m->pointCount = 1; m->type = kCircles;
    m->localNormal = Vec2(0.0f, 0.0f);
    m->localPoint = P;
    m->points[0].id = cf;
    m->points[0].localPoint = circleB.v[0];
    return;
  }
  float den = Dot(e, e);
  Vec2 P = (1.0f / den) * (u * A + v * B);
  Vec2 d = Q - P;
  float dd = Dot(d, d);
  if (dd > radius * radius) return;
  Vec2 n(-e.y, e.x);
  if (Dot(n, Q - A) < 0.0f) n = Vec2(-n.x, -n.y);
  n.Normalize();
  cf.indexA = 0; cf.typeA = kFace;
  m->pointCount = 1; m->type = kFaceA;
  m->localNormal = n;
  m->localPoint = A;
  m->points[0].id = cf;
  m->points[0].localPoint = circleB.v[0];
}

// b2EPCollider::Collide for an edge without ghost vertices
inline void CollideEdgeAndPolygon(Manifold* m, const Shape& edgeA, const Transform& xfA, const Shape& polygonB, const Transform& xfB) {
  Transform xf = MulT(xfA, xfB);
  Vec2 centroidB = Mul(xf, polygonB.centroid);
  Vec2 v1 = edgeA.v[0], v2 = edgeA.v[1];
  Vec2 edge1 = v2 - v1;
  edge1.Normalize();
  Vec2 normal1(edge1.y, -edge1.x);
  float offset1 = Dot(normal1, centroidB - v1);
  bool front = offset1 >= 0.0f;
  Vec2 normal, lowerLimit, upperLimit;
  if (front) { normal = normal1; lowerLimit = -normal1; upperLimit = -normal1; }
  else { normal = -normal1; lowerLimit = normal1; upperLimit = normal1; }
  int count = polygonB.count;
  Vec2 verts[kMaxPolygonVertices], norms[kMaxPolygonVertices];
  for (int i = 0; i < count; ++i) {
    verts[i] = Mul(xf, polygonB.v[i]);
    norms[i] = Mul(xf.q, polygonB.n[i]);
  }
  const float radius = 2.0f * kPolygonRadius;
  m->pointCount = 0;
  // ComputeEdgeSeparation
  int edgeAxisIndex = front ? 0 : 1;
  (void)edgeAxisIndex;
  float edgeSeparation = FLT_MAX;
  for (int i = 0; i < count; ++i) {
    float s = Dot(normal, verts[i] - v1);
    if (s < edgeSeparation) edgeSeparation = s;
  }
  if (edgeSeparation > radius) return;
  // ComputePolygonSeparation
  int polyType = 0;  // 0 unknown, 2 edgeB
  int polyIndex = -1;
  float polySeparation = -FLT_MAX;
  {
    Vec2 perp(-normal.y, normal.x);
    for (int i = 0; i < count; ++i) {
      Vec2 n = -norms[i];
      float s1 = Dot(n, verts[i] - v1);
      float s2 = Dot(n, verts[i] - v2);
      float s = Min(s1, s2);
      if (s > radius) { polyType = 2; polyIndex = i; polySeparation = s; break; }
      if (Dot(n, perp) >= 0.0f) {
        if (Dot(n - upperLimit, normal) < -kAngularSlop) continue;
      } else {
        if (Dot(n - lowerLimit, normal) < -kAngularSlop) continue;
      }
      if (s > polySeparation) { polyType = 2; polyIndex = i; polySeparation = s; }
    }
  }
  if (polyType != 0 && polySeparation > radius) return;
  const float k_relativeTol = 0.98f, k_absoluteTol = 0.001f;
  bool primaryIsEdge;
  if (polyType == 0) primaryIsEdge = true;
  else if (polySeparation > k_relativeTol * edgeSeparation + k_absoluteTol) primaryIsEdge = false;
  else primaryIsEdge = true;
  ClipVertex ie[2];
  int rf_i1, rf_i2;
  Vec2 rf_v1, rf_v2, rf_normal;
  if (primaryIsEdge) {
    m->type = kFaceA;
    int bestIndex = 0;
    float bestValue = Dot(normal, norms[0]);
    for (int i = 1; i < count; ++i) {
      float value = Dot(normal, norms[i]);
      if (value < bestValue) { bestValue = value; bestIndex = i; }
    }
    int i1 = bestIndex, i2 = i1 + 1 < count ? i1 + 1 : 0;
    ie[0].v = verts[i1];
    ie[0].id.indexA = 0; ie[0].id.indexB = (uint8_t)i1; ie[0].id.typeA = kFace; ie[0].id.typeB = kVertex;
    ie[1].v = verts[i2];
    ie[1].id.indexA = 0; ie[1].id.indexB = (uint8_t)i2; ie[1].id.typeA = kFace; ie[1].id.typeB = kVertex;
    if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = v1; rf_v2 = v2; rf_normal = normal1; }
    else { rf_i1 = 1; rf_i2 = 0; rf_v1 = v2; rf_v2 = v1; rf_normal = -normal1; }
  } else {
    m->type = kFaceB;
    ie[0].v = v1;
    ie[0].id.indexA = 0; ie[0].id.indexB = (uint8_t)polyIndex; ie[0].id.typeA = kVertex; ie[0].id.typeB = kFace;
    ie[1].v = v2;
    ie[1].id.indexA = 0; ie[1].id.indexB = (uint8_t)polyIndex; ie[1].id.typeA = kVertex; ie[1].id.typeB = kFace;
    rf_i1 = polyIndex;
    rf_i2 = rf_i1 + 1 < count ? rf_i1 + 1 : 0;
    rf_v1 = verts[rf_i1]; rf_v2 = verts[rf_i2]; rf_normal = norms[rf_i1];
  }
  Vec2 sideNormal1(rf_normal.y, -rf_normal.x);
  Vec2 sideNormal2 = -sideNormal1;
  float sideOffset1 = Dot(sideNormal1, rf_v1);
  float sideOffset2 = Dot(sideNormal2, rf_v2);
  ClipVertex clipPoints1[2], clipPoints2[2];
  int np = ClipSegmentToLine(clipPoints1, ie, sideNormal1, sideOffset1, rf_i1);
  if (np < 2) return;
  np = ClipSegmentToLine(clipPoints2, clipPoints1, sideNormal2, sideOffset2, rf_i2);
  if (np < 2) return;
  if (primaryIsEdge) { m->localNormal = rf_normal; m->localPoint = rf_v1; }
  else { m->localNormal = polygonB.n[rf_i1]; m->localPoint = polygonB.v[rf_i1]; }
  int pointCount = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = Dot(rf_normal, clipPoints2[i].v - rf_v1);
    if (separation <= radius) {
      ManifoldPoint* cp = m->points + pointCount;
      if (primaryIsEdge) {
        cp->localPoint = MulT(xf, clipPoints2[i].v);
        cp->id = clipPoints2[i].id;
      } else {
        cp->localPoint = clipPoints2[i].v;
        cp->id.typeA = clipPoints2[i].id.typeB;
        cp->id.typeB = clipPoints2[i].id.typeA;
        cp->id.indexA = clipPoints2[i].id.indexB;
        cp->id.indexB = clipPoints2[i].id.indexA;
      }
      ++pointCount;
    }
  }
  m->pointCount = pointCount;
}

// ---------------------------------------------------------------------------------------------------------------------
// b2Distance (GJK) and b2TimeOfImpact
struct DistanceProxy {
  const Vec2* vertices = nullptr;
  int count = 0;
  float radius = 0.0f;
  void Set(const Shape& s) { vertices = s.v; count = s.count; radius = s.radius; }
  int GetSupport(const Vec2& d) const {
    int bestIndex = 0;
    float bestValue = Dot(vertices[0], d);
    for (int i = 1; i < count; ++i) {
      float value = Dot(vertices[i], d);
      if (value > bestValue) { bestIndex = i; bestValue = value; }
    }
    return bestIndex;
  }
  const Vec2& GetVertex(int i) const { return vertices[i]; }
};

struct SimplexCache {
  float metric = 0.0f;
  int count = 0;
  int indexA[3] = {0, 0, 0}, indexB[3] = {0, 0, 0};
};

struct SimplexVertex {
  Vec2 wA, wB, w;
  float a = 0.0f;
  int indexA = 0, indexB = 0;
};

struct Simplex {
  SimplexVertex v[3];
  int count = 0;

  float GetMetric() const {
    switch (count) {
      case 1: return 0.0f;
      case 2: return Distance(v[0].w, v[1].w);
      case 3: return Cross(v[1].w - v[0].w, v[2].w - v[0].w);
      default: return 0.0f;
    }
  }
  void ReadCache(const SimplexCache* cache, const DistanceProxy* proxyA, const Transform& xfA, const DistanceProxy* proxyB, const Transform& xfB) {
    count = cache->count;
    for (int i = 0; i < count; ++i) {
      SimplexVertex* sv = v + i;
      sv->indexA = cache->indexA[i];
      sv->indexB = cache->indexB[i];
      sv->wA = Mul(xfA, proxyA->GetVertex(sv->indexA));
      sv->wB = Mul(xfB, proxyB->GetVertex(sv->indexB));
      sv->w = sv->wB - sv->wA;
      sv->a = 0.0f;
    }
    if (count > 1) {
      float metric1 = cache->metric, metric2 = GetMetric();
      if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
    }
    if (count == 0) {
      SimplexVertex* sv = v + 0;
      sv->indexA = 0; sv->indexB = 0;
      sv->wA = Mul(xfA, proxyA->GetVertex(0));
      sv->wB = Mul(xfB, proxyB->GetVertex(0));
      sv->w = sv->wB - sv->wA;
      sv->a = 1.0f;
      count = 1;
    }
  }
  void WriteCache(SimplexCache* cache) const {
    cache->metric = GetMetric();
    cache->count = count;
    for (int i = 0; i < count; ++i) { cache->indexA[i] = v[i].indexA; cache->indexB[i] = v[i].indexB; }
  }
  Vec2 GetSearchDirection() const {
    if (count == 1) return -v[0].w;
    Vec2 e12 = v[1].w - v[0].w;
    float sgn = Cross(e12, -v[0].w);
    if (sgn > 0.0f) return Cross(1.0f, e12);
    return Cross(e12, 1.0f);
  }
  Vec2 GetClosestPoint() const {
    if (count == 1) return v[0].w;
    if (count == 2) return v[0].a * v[0].w + v[1].a * v[1].w;
    return Vec2(0.0f, 0.0f);
  }
  void GetWitnessPoints(Vec2* pA, Vec2* pB) const {
    if (count == 1) { *pA = v[0].wA; *pB = v[0].wB; }
    else if (count == 2) {
      *pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
      *pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
    } else {
      *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
      *pB = *pA;
    }
  }
  void Solve2() {
    Vec2 w1 = v[0].w, w2 = v[1].w;
    Vec2 e12 = w2 - w1;
    float d12_2 = -Dot(w1, e12);
    if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
    float d12_1 = Dot(w2, e12);
    if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
    float inv_d12 = 1.0f / (d12_1 + d12_2);
    v[0].a = d12_1 * inv_d12;
    v[1].a = d12_2 * inv_d12;
    count = 2;
  }
  void Solve3() {
    Vec2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
    Vec2 e12 = w2 - w1;
    float w1e12 = Dot(w1, e12), w2e12 = Dot(w2, e12);
    float d12_1 = w2e12, d12_2 = -w1e12;
    Vec2 e13 = w3 - w1;
    float w1e13 = Dot(w1, e13), w3e13 = Dot(w3, e13);
    float d13_1 = w3e13, d13_2 = -w1e13;
    Vec2 e23 = w3 - w2;
    float w2e23 = Dot(w2, e23), w3e23 = Dot(w3, e23);
    float d23_1 = w3e23, d23_2 = -w2e23;
    float n123 = Cross(e12, e13);
    float d123_1 = n123 * Cross(w2, w3);
    float d123_2 = n123 * Cross(w3, w1);
    float d123_3 = n123 * Cross(w1, w2);
    if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
    if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
      float inv_d12 = 1.0f / (d12_1 + d12_2);
      v[0].a = d12_1 * inv_d12; v[1].a = d12_2 * inv_d12; count = 2;
      return;
    }
    if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
      float inv_d13 = 1.0f / (d13_1 + d13_2);
      v[0].a = d13_1 * inv_d13; v[2].a = d13_2 * inv_d13; count = 2; v[1] = v[2];
      return;
    }
    if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
    if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; return; }
    if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
      float inv_d23 = 1.0f / (d23_1 + d23_2);
      v[1].a = d23_1 * inv_d23; v[2].a = d23_2 * inv_d23; count = 2; v[0] = v[2];
      return;
    }
    float inv_d123 = 1.0f / (d123_1 + d123_2 + d123_3);
    v[0].a = d123_1 * inv_d123; v[1].a = d123_2 * inv_d123; v[2].a = d123_3 * inv_d123;
    count = 3;
  }
};

struct DistanceOutput {
  Vec2 pointA, pointB;
  float distance = 0.0f;
  int iterations = 0;
};

// b2Distance with useRadii = false (the only use on this path: inside b2TimeOfImpact)
inline void DistanceGJK(DistanceOutput* output, SimplexCache* cache, const DistanceProxy* proxyA, const Transform& xfA,
                        const DistanceProxy* proxyB, const Transform& xfB) {
  Simplex simplex;
  simplex.ReadCache(cache, proxyA, xfA, proxyB, xfB);
  SimplexVertex* vertices = simplex.v;
  const int k_maxIters = 20;
  int saveA[3], saveB[3];
  int saveCount = 0;
  int iter = 0;
  while (iter < k_maxIters) {
    saveCount = simplex.count;
    for (int i = 0; i < saveCount; ++i) { saveA[i] = vertices[i].indexA; saveB[i] = vertices[i].indexB; }
    switch (simplex.count) {
      case 1: break;
      case 2: simplex.Solve2(); break;
      case 3: simplex.Solve3(); break;
    }
    if (simplex.count == 3) break;
    Vec2 d = simplex.GetSearchDirection();
    if (d.LengthSquared() < kEpsilon * kEpsilon) break;
    SimplexVertex* vertex = vertices + simplex.count;
    vertex->indexA = proxyA->GetSupport(MulT(xfA.q, -d));
    vertex->wA = Mul(xfA, proxyA->GetVertex(vertex->indexA));
    vertex->indexB = proxyB->GetSupport(MulT(xfB.q, d));
    vertex->wB = Mul(xfB, proxyB->GetVertex(vertex->indexB));
    vertex->w = vertex->wB - vertex->wA;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < saveCount; ++i) {
      if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) { duplicate = true; break; }
    }
    if (duplicate) break;
    ++simplex.count;
  }
  simplex.GetWitnessPoints(&output->pointA, &output->pointB);
  output->distance = Distance(output->pointA, output->pointB);
  output->iterations = iter;
  simplex.WriteCache(cache);
}

struct SeparationFunction {
  enum Type { kPoints, kFaceA_, kFaceB_ };
  const DistanceProxy *proxyA, *proxyB;
  Sweep sweepA, sweepB;
  Type type;
  Vec2 localPoint, axis;

  float Initialize(const SimplexCache* cache, const DistanceProxy* pA, const Sweep& sA, const DistanceProxy* pB, const Sweep& sB, float t1) {
    proxyA = pA; proxyB = pB;
    int count = cache->count;
    sweepA = sA; sweepB = sB;
    Transform xfA, xfB;
    sweepA.GetTransform(&xfA, t1);
    sweepB.GetTransform(&xfB, t1);
    if (count == 1) {
      type = kPoints;
      Vec2 pointA = Mul(xfA, proxyA->GetVertex(cache->indexA[0]));
      Vec2 pointB = Mul(xfB, proxyB->GetVertex(cache->indexB[0]));
      axis = pointB - pointA;
      return axis.Normalize();
    } else if (cache->indexA[0] == cache->indexA[1]) {
      type = kFaceB_;
      Vec2 localPointB1 = proxyB->GetVertex(cache->indexB[0]);
      Vec2 localPointB2 = proxyB->GetVertex(cache->indexB[1]);
      axis = Cross(localPointB2 - localPointB1, 1.0f);
      axis.Normalize();
      Vec2 normal = Mul(xfB.q, axis);
      localPoint = 0.5f * (localPointB1 + localPointB2);
      Vec2 pointB = Mul(xfB, localPoint);
      Vec2 pointA = Mul(xfA, proxyA->GetVertex(cache->indexA[0]));
      float s = Dot(pointA - pointB, normal);
      if (s < 0.0f) { axis = -axis; s = -s; }
      return s;
    } else {
      type = kFaceA_;
      Vec2 localPointA1 = proxyA->GetVertex(cache->indexA[0]);
      Vec2 localPointA2 = proxyA->GetVertex(cache->indexA[1]);
      axis = Cross(localPointA2 - localPointA1, 1.0f);
      axis.Normalize();
      Vec2 normal = Mul(xfA.q, axis);
      localPoint = 0.5f * (localPointA1 + localPointA2);
      Vec2 pointA = Mul(xfA, localPoint);
      Vec2 pointB = Mul(xfB, proxyB->GetVertex(cache->indexB[0]));
      float s = Dot(pointB - pointA, normal);
      if (s < 0.0f) { axis = -axis; s = -s; }
      return s;
    }
  }
  float FindMinSeparation(int* indexA, int* indexB, float t) const {
    Transform xfA, xfB;
    sweepA.GetTransform(&xfA, t);
    sweepB.GetTransform(&xfB, t);
    switch (type) {
      case kPoints: {
        Vec2 axisA = MulT(xfA.q, axis);
        Vec2 axisB = MulT(xfB.q, -axis);
        *indexA = proxyA->GetSupport(axisA);
        *indexB = proxyB->GetSupport(axisB);
        Vec2 pointA = Mul(xfA, proxyA->GetVertex(*indexA));
        Vec2 pointB = Mul(xfB, proxyB->GetVertex(*indexB));
        return Dot(pointB - pointA, axis);
      }
      case kFaceA_: {
        Vec2 normal = Mul(xfA.q, axis);
        Vec2 pointA = Mul(xfA, localPoint);
        Vec2 axisB = MulT(xfB.q, -normal);
        *indexA = -1;
        *indexB = proxyB->GetSupport(axisB);
        Vec2 pointB = Mul(xfB, proxyB->GetVertex(*indexB));
        return Dot(pointB - pointA, normal);
      }
      default: {
        Vec2 normal = Mul(xfB.q, axis);
        Vec2 pointB = Mul(xfB, localPoint);
        Vec2 axisA = MulT(xfA.q, -normal);
        *indexB = -1;
        *indexA = proxyA->GetSupport(axisA);
        Vec2 pointA = Mul(xfA, proxyA->GetVertex(*indexA));
        return Dot(pointA - pointB, normal);
      }
    }
  }
  float Evaluate(int indexA, int indexB, float t) const {
    Transform xfA, xfB;
    sweepA.GetTransform(&xfA, t);
    sweepB.GetTransform(&xfB, t);
    switch (type) {
      case kPoints: {
        Vec2 pointA = Mul(xfA, proxyA->GetVertex(indexA));
        Vec2 pointB = Mul(xfB, proxyB->GetVertex(indexB));
        return Dot(pointB - pointA, axis);
      }
      case kFaceA_: {
        Vec2 normal = Mul(xfA.q, axis);
        Vec2 pointA = Mul(xfA, localPoint);
        Vec2 pointB = Mul(xfB, proxyB->GetVertex(indexB));
        return Dot(pointB - pointA, normal);
      }
      default: {
        Vec2 normal = Mul(xfB.q, axis);
        Vec2 pointB = Mul(xfB, localPoint);
        Vec2 pointA = Mul(xfA, proxyA->GetVertex(indexA));
        return Dot(pointA - pointB, normal);
      }
    }
  }
};

enum ToiState { kToiUnknown, kToiFailed, kToiOverlapped, kToiTouching, kToiSeparated };

struct ToiOutput {
  ToiState state = kToiUnknown;
  float t = 0.0f;
};

inline void TimeOfImpact(ToiOutput* output, const DistanceProxy* proxyA, const DistanceProxy* proxyB, Sweep sweepA, Sweep sweepB, float tMax) {
  output->state = kToiUnknown;
  output->t = tMax;
  sweepA.Normalize();
  sweepB.Normalize();
  float totalRadius = proxyA->radius + proxyB->radius;
  float target = Max(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
  float tolerance = 0.25f * kLinearSlop;
  float t1 = 0.0f;
  const int k_maxIterations = 20;
  int iter = 0;
  SimplexCache cache;
  cache.count = 0;
  for (;;) {
    Transform xfA, xfB;
    sweepA.GetTransform(&xfA, t1);
    sweepB.GetTransform(&xfB, t1);
    DistanceOutput distanceOutput;
    DistanceGJK(&distanceOutput, &cache, proxyA, xfA, proxyB, xfB);
    if (distanceOutput.distance <= 0.0f) { output->state = kToiOverlapped; output->t = 0.0f; break; }
    if (distanceOutput.distance < target + tolerance) { output->state = kToiTouching; output->t = t1; break; }
    SeparationFunction fcn;
    fcn.Initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1);
    bool done = false;
    float t2 = tMax;
    int pushBackIter = 0;
    for (;;) {
      int indexA, indexB;
      float s2 = fcn.FindMinSeparation(&indexA, &indexB, t2);
      if (s2 > target + tolerance) { output->state = kToiSeparated; output->t = tMax; done = true; break; }
      if (s2 > target - tolerance) { t1 = t2; break; }
      float s1 = fcn.Evaluate(indexA, indexB, t1);
      if (s1 < target - tolerance) { output->state = kToiFailed; output->t = t1; done = true; break; }
      if (s1 <= target + tolerance) { output->state = kToiTouching; output->t = t1; done = true; break; }
      int rootIterCount = 0;
      float a1 = t1, a2 = t2;
      for (;;) {
        float t;
        if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
        else t = 0.5f * (a1 + a2);
        ++rootIterCount;
        float s = fcn.Evaluate(indexA, indexB, t);
        if (Abs(s - target) < tolerance) { t2 = t; break; }
        if (s > target) { a1 = t; s1 = s; }
        else { a2 = t; s2 = s; }
        if (rootIterCount == 50) break;
      }
      ++pushBackIter;
      if (pushBackIter == kMaxPolygonVertices) break;
    }
    ++iter;
    if (done) break;
    if (iter == k_maxIterations) { output->state = kToiFailed; output->t = t1; break; }
  }
}

}  // namespace b2o
