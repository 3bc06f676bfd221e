// blcd_collide.cuh -- narrow phase, GJK distance and time of impact for one pair of fixtures (host+device).
//
// Algorithms of Box2D 2.3.x b2CollideCircle.cpp / b2CollidePolygon.cpp / b2CollideEdge.cpp / b2Collision.cpp /
// b2Distance.cpp / b2TimeOfImpact.cpp, the C++ the reference reaches through b2World.Step
// (boxLCD/world_env.py:446-452), on the fixed-size scene tables of blcd_scene.h.  Manifolds are plain register
// structs; contact ids are packed into one word.  Static fixtures (the arena walls) sit at the identity transform.
#pragma once
#include "blcd_scene.h"

namespace BLCD_NS {

enum { MF_CIRCLES = 0, MF_FACE_A = 1, MF_FACE_B = 2 };
enum { FT_VERTEX = 0, FT_FACE = 1 };

BLCD_HD uint32_t mkid(int ia, int ib, int ta, int tb) { return (uint32_t)ia | ((uint32_t)ib << 8) | ((uint32_t)ta << 16) | ((uint32_t)tb << 24); }
BLCD_HD uint32_t id_swap(uint32_t id) { return ((id >> 8) & 0xFFu) | ((id & 0xFFu) << 8) | (((id >> 24) & 0xFFu) << 16) | (((id >> 16) & 0xFFu) << 24); }

struct Mf {  // b2Manifold
  int type, count;
  V2 ln, lp;
  V2 pt[2];
  float ni[2], ti[2];
  uint32_t id[2];
};

struct CV {  // b2ClipVertex
  V2 v;
  uint32_t id;
};

BLCD_HD void collide_circles(Mf& m, const DShape& A, const Xf& xfA, const DShape& B, const Xf& xfB) {
  m.count = 0;
  V2 pA = xmul(xfA, A.v[0]), pB = xmul(xfB, B.v[0]);
  V2 d = pB - pA;
  float distSqr = dot(d, d);
  float radius = A.radius + B.radius;
  if (distSqr > radius * radius) return;
  m.type = MF_CIRCLES;
  m.lp = A.v[0];
  m.ln = mk(0.0f, 0.0f);
  m.count = 1;
  m.pt[0] = B.v[0];
  m.id[0] = 0u;
}

BLCD_HDN void collide_polygon_circle(Mf& m, const DShape& A, const Xf& xfA, const DShape& B, const Xf& xfB) {
  m.count = 0;
  V2 c = xmul(xfB, B.v[0]);
  V2 cLocal = xmulT(xfA, c);
  int normalIndex = 0;
  float separation = -kMaxFloat;
  float radius = A.radius + B.radius;
  for (int i = 0; i < A.count; ++i) {
    float s = dot(A.n[i], cLocal - A.v[i]);
    if (s > radius) return;
    if (s > separation) { separation = s; normalIndex = i; }
  }
  int i1 = normalIndex, i2 = i1 + 1 < A.count ? i1 + 1 : 0;
  V2 v1 = A.v[i1], v2 = A.v[i2];
  m.pt[0] = B.v[0];
  m.id[0] = 0u;
  m.type = MF_FACE_A;
  if (separation < kEps) {
    m.count = 1;
    m.ln = A.n[normalIndex];
    m.lp = 0.5f * (v1 + v2);
    return;
  }
  float u1 = dot(cLocal - v1, v2 - v1);
  float u2 = dot(cLocal - v2, v1 - v2);
  if (u1 <= 0.0f) {
    if (dist2(cLocal, v1) > radius * radius) return;
    m.count = 1;
    m.ln = cLocal - v1;
    normalize(m.ln);
    m.lp = v1;
  } else if (u2 <= 0.0f) {
    if (dist2(cLocal, v2) > radius * radius) return;
    m.count = 1;
    m.ln = cLocal - v2;
    normalize(m.ln);
    m.lp = v2;
  } else {
    V2 faceCenter = 0.5f * (v1 + v2);
    float sep = dot(cLocal - faceCenter, A.n[i1]);
    if (sep > radius) return;
    m.count = 1;
    m.ln = A.n[i1];
    m.lp = faceCenter;
  }
}

BLCD_HD int clip_segment(CV out[2], const CV in[2], V2 normal, float offset, int vertexIndexA) {
  int numOut = 0;
  float d0 = dot(normal, in[0].v) - offset;
  float d1 = dot(normal, in[1].v) - offset;
  if (d0 <= 0.0f) out[numOut++] = in[0];
  if (d1 <= 0.0f) out[numOut++] = in[1];
  if (d0 * d1 < 0.0f) {
    float interp = d0 / (d0 - d1);
    out[numOut].v = in[0].v + interp * (in[1].v - in[0].v);
    out[numOut].id = mkid(vertexIndexA, (int)((in[0].id >> 8) & 0xFFu), FT_VERTEX, FT_FACE);
    ++numOut;
  }
  return numOut;
}

BLCD_HD float find_max_separation(int* edgeIndex, const DShape& p1, const Xf& xf1, const DShape& p2, const Xf& xf2) {
  Xf xf = xmulT(xf2, xf1);
  int best = 0;
  float maxSep = -kMaxFloat;
  for (int i = 0; i < p1.count; ++i) {
    V2 n = rmul(xf.q, p1.n[i]);
    V2 v1 = xmul(xf, p1.v[i]);
    float si = kMaxFloat;
    for (int j = 0; j < p2.count; ++j) {
      float sij = dot(n, p2.v[j] - v1);
      if (sij < si) si = sij;
    }
    if (si > maxSep) { maxSep = si; best = i; }
  }
  *edgeIndex = best;
  return maxSep;
}

// Box2D 2.3.0's b2EdgeSeparation / hill-climbing b2FindMaxSeparation (the brute-force loop above is 2.3.1's).  Selected
// together with the 2.3.0 reference-face hysteresis by BLCD_FLAG_REFFACE_2_3_0: the reference's recorded Object2-cubes
// episode is reproduced with these rules, i.e. this is what pybox2d 2.3.10 vendors.
BLCD_HD float edge_separation_230(const DShape& p1, const Xf& xf1, int edge1, const DShape& p2, const Xf& xf2) {
  V2 normal1World = rmul(xf1.q, p1.n[edge1]);
  V2 normal1 = rmulT(xf2.q, normal1World);
  int index = 0;
  float minDot = kMaxFloat;
  for (int i = 0; i < p2.count; ++i) {
    float d = dot(p2.v[i], normal1);
    if (d < minDot) { minDot = d; index = i; }
  }
  V2 v1 = xmul(xf1, p1.v[edge1]);
  V2 v2 = xmul(xf2, p2.v[index]);
  return dot(v2 - v1, normal1World);
}

BLCD_HD float find_max_separation_230(int* edgeIndex, const DShape& p1, const Xf& xf1, const DShape& p2, const Xf& xf2) {
  int count1 = p1.count;
  V2 d = xmul(xf2, p2.centroid) - xmul(xf1, p1.centroid);
  V2 dLocal1 = rmulT(xf1.q, d);
  int edge = 0;
  float maxDot = -kMaxFloat;
  for (int i = 0; i < count1; ++i) {
    float dt = dot(p1.n[i], dLocal1);
    if (dt > maxDot) { maxDot = dt; edge = i; }
  }
  float s = edge_separation_230(p1, xf1, edge, p2, xf2);
  int prevEdge = edge - 1 >= 0 ? edge - 1 : count1 - 1;
  float sPrev = edge_separation_230(p1, xf1, prevEdge, p2, xf2);
  int nextEdge = edge + 1 < count1 ? edge + 1 : 0;
  float sNext = edge_separation_230(p1, xf1, nextEdge, p2, xf2);
  int bestEdge, increment;
  float bestSeparation;
  if (sPrev > s && sPrev > sNext) { increment = -1; bestEdge = prevEdge; bestSeparation = sPrev; }
  else if (sNext > s) { increment = 1; bestEdge = nextEdge; bestSeparation = sNext; }
  else { *edgeIndex = edge; return s; }
  for (;;) {
    if (increment == -1) edge = bestEdge - 1 >= 0 ? bestEdge - 1 : count1 - 1;
    else edge = bestEdge + 1 < count1 ? bestEdge + 1 : 0;
    s = edge_separation_230(p1, xf1, edge, p2, xf2);
    if (s > bestSeparation) { bestEdge = edge; bestSeparation = s; }
    else break;
  }
  *edgeIndex = bestEdge;
  return bestSeparation;
}

BLCD_HDN void collide_polygons(Mf& m, const DShape& polyA, const Xf& xfA, const DShape& polyB, const Xf& xfB, bool refface_2_3_0) {
  m.count = 0;
  float totalRadius = polyA.radius + polyB.radius;
  int edgeA = 0;
  float sepA = refface_2_3_0 ? find_max_separation_230(&edgeA, polyA, xfA, polyB, xfB) : find_max_separation(&edgeA, polyA, xfA, polyB, xfB);
  if (sepA > totalRadius) return;
  int edgeB = 0;
  float sepB = refface_2_3_0 ? find_max_separation_230(&edgeB, polyB, xfB, polyA, xfA) : find_max_separation(&edgeB, polyB, xfB, polyA, xfA);
  if (sepB > totalRadius) return;
  bool useB = refface_2_3_0 ? (sepB > 0.98f * sepA + 0.001f) : (sepB > sepA + 0.1f * kLinearSlop);
  const DShape& poly1 = useB ? polyB : polyA;
  const DShape& poly2 = useB ? polyA : polyB;
  const Xf xf1 = useB ? xfB : xfA, xf2 = useB ? xfA : xfB;
  const int edge1 = useB ? edgeB : edgeA;
  m.type = useB ? MF_FACE_B : MF_FACE_A;
  // b2FindIncidentEdge
  CV incident[2];
  {
    V2 normal1 = rmulT(xf2.q, rmul(xf1.q, poly1.n[edge1]));
    int index = 0;
    float minDot = kMaxFloat;
    for (int i = 0; i < poly2.count; ++i) {
      float d = dot(normal1, poly2.n[i]);
      if (d < minDot) { minDot = d; index = i; }
    }
    int i1 = index, i2 = i1 + 1 < poly2.count ? i1 + 1 : 0;
    incident[0].v = xmul(xf2, poly2.v[i1]);
    incident[0].id = mkid(edge1, i1, FT_FACE, FT_VERTEX);
    incident[1].v = xmul(xf2, poly2.v[i2]);
    incident[1].id = mkid(edge1, i2, FT_FACE, FT_VERTEX);
  }
  int iv1 = edge1, iv2 = edge1 + 1 < poly1.count ? edge1 + 1 : 0;
  V2 v11 = poly1.v[iv1], v12 = poly1.v[iv2];
  V2 localTangent = v12 - v11;
  normalize(localTangent);
  V2 localNormal = cross(localTangent, 1.0f);
  V2 planePoint = 0.5f * (v11 + v12);
  V2 tangent = rmul(xf1.q, localTangent);
  V2 normal = cross(tangent, 1.0f);
  v11 = xmul(xf1, v11);
  v12 = xmul(xf1, v12);
  float frontOffset = dot(normal, v11);
  float sideOffset1 = -dot(tangent, v11) + totalRadius;
  float sideOffset2 = dot(tangent, v12) + totalRadius;
  CV cp1[2], cp2[2];
  int np = clip_segment(cp1, incident, -tangent, sideOffset1, iv1);
  if (np < 2) return;
  np = clip_segment(cp2, cp1, tangent, sideOffset2, iv2);
  if (np < 2) return;
  m.ln = localNormal;
  m.lp = planePoint;
  int pc = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = dot(normal, cp2[i].v) - frontOffset;
    if (separation <= totalRadius) {
      m.pt[pc] = xmulT(xf2, cp2[i].v);
      m.id[pc] = useB ? id_swap(cp2[i].id) : cp2[i].id;
      ++pc;
    }
  }
  m.count = pc;
}

// edge A is a wall: static body at the identity transform, no ghost vertices (world_env.py:311-316)
BLCD_HDN void collide_edge_circle(Mf& m, const DShape& edgeA, const DShape& circleB, const Xf& xfB) {
  m.count = 0;
  V2 Q = xmul(xfB, circleB.v[0]);  // MulT by the identity is exact
  V2 A = edgeA.v[0], B = edgeA.v[1];
  V2 e = B - A;
  float u = dot(e, B - Q);
  float v = dot(e, Q - A);
  float radius = edgeA.radius + circleB.radius;
  m.pt[0] = circleB.v[0];
  if (v <= 0.0f) {
    V2 d = Q - A;
    if (dot(d, d) > radius * radius) return;
    m.count = 1; m.type = MF_CIRCLES; m.ln = mk(0.0f, 0.0f); m.lp = A;
    m.id[0] = mkid(0, 0, FT_VERTEX, FT_VERTEX);
    return;
  }
  if (u <= 0.0f) {
    V2 d = Q - B;
    if (dot(d, d) > radius * radius) return;
    m.count = 1; m.type = MF_CIRCLES; m.ln = mk(0.0f, 0.0f); m.lp = B;
    m.id[0] = mkid(1, 0, FT_VERTEX, FT_VERTEX);
    return;
  }
  float den = dot(e, e);
  V2 P = (1.0f / den) * (u * A + v * B);
  V2 d = Q - P;
  if (dot(d, d) > radius * radius) return;
  V2 n = mk(-e.y, e.x);
  if (dot(n, Q - A) < 0.0f) n = mk(-n.x, -n.y);
  normalize(n);
  m.count = 1; m.type = MF_FACE_A; m.ln = n; m.lp = A;
  m.id[0] = mkid(0, 0, FT_FACE, FT_VERTEX);
}

// b2EPCollider::Collide, edge A at the identity transform and without ghost vertices
BLCD_HDN void collide_edge_polygon(Mf& m, const DShape& edgeA, const DShape& polyB, const Xf& xfB) {
  m.count = 0;
  const Xf xf = xfB;  // MulT(identity, xfB)
  V2 centroidB = xmul(xf, polyB.centroid);
  V2 v1 = edgeA.v[0], v2 = edgeA.v[1];
  V2 edge1 = v2 - v1;
  normalize(edge1);
  V2 normal1 = mk(edge1.y, -edge1.x);
  float offset1 = dot(normal1, centroidB - v1);
  bool front = offset1 >= 0.0f;
  V2 normal = front ? normal1 : -normal1;
  V2 limit = front ? -normal1 : normal1;  // lower == upper limit without adjacent edges
  const int count = polyB.count;
  V2 verts[BLCD_MAX_VERTS], norms[BLCD_MAX_VERTS];
  for (int i = 0; i < count; ++i) {
    verts[i] = xmul(xf, polyB.v[i]);
    norms[i] = rmul(xf.q, polyB.n[i]);
  }
  const float radius = 2.0f * kPolygonRadius;
  float edgeSep = FLT_MAX;
  for (int i = 0; i < count; ++i) {
    float s = dot(normal, verts[i] - v1);
    if (s < edgeSep) edgeSep = s;
  }
  if (edgeSep > radius) return;
  int polyIndex = -1;
  float polySep = -FLT_MAX;
  {
    V2 perp = mk(-normal.y, normal.x);
    for (int i = 0; i < count; ++i) {
      V2 n = -norms[i];
      float s1 = dot(n, verts[i] - v1);
      float s2 = dot(n, verts[i] - v2);
      float s = fminb(s1, s2);
      if (s > radius) { polyIndex = i; polySep = s; break; }
      if (dot(n, perp) >= 0.0f) {
        if (dot(n - limit, normal) < -kAngularSlop) continue;
      } else {
        if (dot(n - limit, normal) < -kAngularSlop) continue;
      }
      if (s > polySep) { polyIndex = i; polySep = s; }
    }
  }
  if (polyIndex >= 0 && polySep > radius) return;
  bool primaryIsEdge = !(polyIndex >= 0 && polySep > 0.98f * edgeSep + 0.001f);
  CV ie[2];
  int rf_i1, rf_i2;
  V2 rf_v1, rf_v2, rf_normal;
  if (primaryIsEdge) {
    m.type = MF_FACE_A;
    int best = 0;
    float bestValue = dot(normal, norms[0]);
    for (int i = 1; i < count; ++i) {
      float value = dot(normal, norms[i]);
      if (value < bestValue) { bestValue = value; best = i; }
    }
    int i1 = best, i2 = i1 + 1 < count ? i1 + 1 : 0;
    ie[0].v = verts[i1]; ie[0].id = mkid(0, i1, FT_FACE, FT_VERTEX);
    ie[1].v = verts[i2]; ie[1].id = mkid(0, i2, FT_FACE, FT_VERTEX);
    if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = v1; rf_v2 = v2; rf_normal = normal1; }
    else { rf_i1 = 1; rf_i2 = 0; rf_v1 = v2; rf_v2 = v1; rf_normal = -normal1; }
  } else {
    m.type = MF_FACE_B;
    ie[0].v = v1; ie[0].id = mkid(0, polyIndex, FT_VERTEX, FT_FACE);
    ie[1].v = v2; ie[1].id = mkid(0, polyIndex, FT_VERTEX, FT_FACE);
    rf_i1 = polyIndex;
    rf_i2 = rf_i1 + 1 < count ? rf_i1 + 1 : 0;
    rf_v1 = verts[rf_i1]; rf_v2 = verts[rf_i2]; rf_normal = norms[rf_i1];
  }
  V2 sideNormal1 = mk(rf_normal.y, -rf_normal.x);
  V2 sideNormal2 = -sideNormal1;
  float sideOffset1 = dot(sideNormal1, rf_v1);
  float sideOffset2 = dot(sideNormal2, rf_v2);
  CV cp1[2], cp2[2];
  int np = clip_segment(cp1, ie, sideNormal1, sideOffset1, rf_i1);
  if (np < 2) return;
  np = clip_segment(cp2, cp1, sideNormal2, sideOffset2, rf_i2);
  if (np < 2) return;
  if (primaryIsEdge) { m.ln = rf_normal; m.lp = rf_v1; }
  else { m.ln = polyB.n[rf_i1]; m.lp = polyB.v[rf_i1]; }
  int pc = 0;
  for (int i = 0; i < 2; ++i) {
    float separation = dot(rf_normal, cp2[i].v - rf_v1);
    if (separation <= radius) {
      if (primaryIsEdge) {
        m.pt[pc] = xmulT(xf, cp2[i].v);
        m.id[pc] = cp2[i].id;
      } else {
        m.pt[pc] = cp2[i].v;
        m.id[pc] = id_swap(cp2[i].id);
      }
      ++pc;
    }
  }
  m.count = pc;
}

// ---------------------------------------------------------------------------------------------------------------------
// GJK distance between a wall edge (identity transform) or any convex fixture and a moving fixture, and the
// conservative-advancement time of impact built on it.
BLCD_HD int support(const DShape& s, V2 d) {
  int best = 0;
  float bestValue = dot(s.v[0], d);
  for (int i = 1; i < s.count; ++i) {
    float value = dot(s.v[i], d);
    if (value > bestValue) { best = i; bestValue = value; }
  }
  return best;
}

struct SimplexCache {
  float metric;
  int count;
  int ia[3], ib[3];
};

struct SVert {
  V2 wA, wB, w;
  float a;
  int ia, ib;
};

struct Simplex {
  SVert v[3];
  int count;
};

BLCD_HD float simplex_metric(const Simplex& s) {
  if (s.count == 2) return len(s.v[0].w - s.v[1].w);
  if (s.count == 3) return cross(s.v[1].w - s.v[0].w, s.v[2].w - s.v[0].w);
  return 0.0f;
}

BLCD_HD void simplex_solve2(Simplex& s) {
  V2 w1 = s.v[0].w, w2 = s.v[1].w;
  V2 e12 = w2 - w1;
  float d12_2 = -dot(w1, e12);
  if (d12_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
  float d12_1 = dot(w2, e12);
  if (d12_1 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
  float inv = 1.0f / (d12_1 + d12_2);
  s.v[0].a = d12_1 * inv;
  s.v[1].a = d12_2 * inv;
  s.count = 2;
}

BLCD_HD void simplex_solve3(Simplex& s) {
  V2 w1 = s.v[0].w, w2 = s.v[1].w, w3 = s.v[2].w;
  V2 e12 = w2 - w1;
  float w1e12 = dot(w1, e12), w2e12 = dot(w2, e12);
  float d12_1 = w2e12, d12_2 = -w1e12;
  V2 e13 = w3 - w1;
  float w1e13 = dot(w1, e13), w3e13 = dot(w3, e13);
  float d13_1 = w3e13, d13_2 = -w1e13;
  V2 e23 = w3 - w2;
  float w2e23 = dot(w2, e23), w3e23 = dot(w3, e23);
  float d23_1 = w3e23, d23_2 = -w2e23;
  float n123 = cross(e12, e13);
  float d123_1 = n123 * cross(w2, w3);
  float d123_2 = n123 * cross(w3, w1);
  float d123_3 = n123 * cross(w1, w2);
  if (d12_2 <= 0.0f && d13_2 <= 0.0f) { s.v[0].a = 1.0f; s.count = 1; return; }
  if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
    float inv = 1.0f / (d12_1 + d12_2);
    s.v[0].a = d12_1 * inv; s.v[1].a = d12_2 * inv; s.count = 2;
    return;
  }
  if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
    float inv = 1.0f / (d13_1 + d13_2);
    s.v[0].a = d13_1 * inv; s.v[2].a = d13_2 * inv; s.count = 2; s.v[1] = s.v[2];
    return;
  }
  if (d12_1 <= 0.0f && d23_2 <= 0.0f) { s.v[1].a = 1.0f; s.count = 1; s.v[0] = s.v[1]; return; }
  if (d13_1 <= 0.0f && d23_1 <= 0.0f) { s.v[2].a = 1.0f; s.count = 1; s.v[0] = s.v[2]; return; }
  if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
    float inv = 1.0f / (d23_1 + d23_2);
    s.v[1].a = d23_1 * inv; s.v[2].a = d23_2 * inv; s.count = 2; s.v[0] = s.v[2];
    return;
  }
  float inv = 1.0f / (d123_1 + d123_2 + d123_3);
  s.v[0].a = d123_1 * inv; s.v[1].a = d123_2 * inv; s.v[2].a = d123_3 * inv;
  s.count = 3;
}

// b2Distance with useRadii = false; returns the distance between the cores
BLCD_HDN float gjk_distance(SimplexCache& cache, const DShape& A, const Xf& xfA, const DShape& B, const Xf& xfB) {
  Simplex sx;
  sx.count = cache.count;
  for (int i = 0; i < sx.count; ++i) {
    SVert& sv = sx.v[i];
    sv.ia = cache.ia[i]; sv.ib = cache.ib[i];
    sv.wA = xmul(xfA, A.v[sv.ia]);
    sv.wB = xmul(xfB, B.v[sv.ib]);
    sv.w = sv.wB - sv.wA;
    sv.a = 0.0f;
  }
  if (sx.count > 1) {
    float metric1 = cache.metric, metric2 = simplex_metric(sx);
    if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEps) sx.count = 0;
  }
  if (sx.count == 0) {
    SVert& sv = sx.v[0];
    sv.ia = 0; sv.ib = 0;
    sv.wA = xmul(xfA, A.v[0]);
    sv.wB = xmul(xfB, B.v[0]);
    sv.w = sv.wB - sv.wA;
    sv.a = 1.0f;
    sx.count = 1;
  }
  int saveA[3], saveB[3];
  int iter = 0;
  while (iter < 20) {
    int saveCount = sx.count;
    for (int i = 0; i < saveCount; ++i) { saveA[i] = sx.v[i].ia; saveB[i] = sx.v[i].ib; }
    if (sx.count == 2) simplex_solve2(sx);
    else if (sx.count == 3) simplex_solve3(sx);
    if (sx.count == 3) break;
    V2 d;
    if (sx.count == 1) d = -sx.v[0].w;
    else {
      V2 e12 = sx.v[1].w - sx.v[0].w;
      float sgn = cross(e12, -sx.v[0].w);
      d = sgn > 0.0f ? cross(1.0f, e12) : cross(e12, 1.0f);
    }
    if (len2(d) < kEps * kEps) break;
    SVert& nv = sx.v[sx.count];
    nv.ia = support(A, rmulT(xfA.q, -d));
    nv.wA = xmul(xfA, A.v[nv.ia]);
    nv.ib = support(B, rmulT(xfB.q, d));
    nv.wB = xmul(xfB, B.v[nv.ib]);
    nv.w = nv.wB - nv.wA;
    ++iter;
    bool duplicate = false;
    for (int i = 0; i < saveCount; ++i)
      if (nv.ia == saveA[i] && nv.ib == saveB[i]) { duplicate = true; break; }
    if (duplicate) break;
    ++sx.count;
  }
  V2 pA, pB;
  if (sx.count == 1) { pA = sx.v[0].wA; pB = sx.v[0].wB; }
  else if (sx.count == 2) {
    pA = sx.v[0].a * sx.v[0].wA + sx.v[1].a * sx.v[1].wA;
    pB = sx.v[0].a * sx.v[0].wB + sx.v[1].a * sx.v[1].wB;
  } else {
    pA = sx.v[0].a * sx.v[0].wA + sx.v[1].a * sx.v[1].wA + sx.v[2].a * sx.v[2].wA;
    pB = pA;
  }
  cache.metric = simplex_metric(sx);
  cache.count = sx.count;
  for (int i = 0; i < sx.count; ++i) { cache.ia[i] = sx.v[i].ia; cache.ib[i] = sx.v[i].ib; }
  return len(pA - pB);
}

struct SepFn {  // b2SeparationFunction
  int type;     // 0 points, 1 faceA, 2 faceB
  V2 localPoint, axis;
};

BLCD_HD void sepfn_init(SepFn& f, const SimplexCache& cache, const DShape& A, const Sweep& sA, const DShape& B, const Sweep& sB, float t1, bool a_static) {
  Xf xfA = a_static ? xf_identity() : sweep_xf(sA, t1), xfB = sweep_xf(sB, t1);
  if (cache.count == 1) {
    f.type = 0;
    V2 pointA = xmul(xfA, A.v[cache.ia[0]]);
    V2 pointB = xmul(xfB, B.v[cache.ib[0]]);
    f.axis = pointB - pointA;
    normalize(f.axis);
  } else if (cache.ia[0] == cache.ia[1]) {
    f.type = 2;
    V2 b1 = B.v[cache.ib[0]], b2 = B.v[cache.ib[1]];
    f.axis = cross(b2 - b1, 1.0f);
    normalize(f.axis);
    V2 normal = rmul(xfB.q, f.axis);
    f.localPoint = 0.5f * (b1 + b2);
    V2 pointB = xmul(xfB, f.localPoint);
    V2 pointA = xmul(xfA, A.v[cache.ia[0]]);
    float s = dot(pointA - pointB, normal);
    if (s < 0.0f) f.axis = -f.axis;
  } else {
    f.type = 1;
    V2 a1 = A.v[cache.ia[0]], a2 = A.v[cache.ia[1]];
    f.axis = cross(a2 - a1, 1.0f);
    normalize(f.axis);
    V2 normal = rmul(xfA.q, f.axis);
    f.localPoint = 0.5f * (a1 + a2);
    V2 pointA = xmul(xfA, f.localPoint);
    V2 pointB = xmul(xfB, B.v[cache.ib[0]]);
    float s = dot(pointB - pointA, normal);
    if (s < 0.0f) f.axis = -f.axis;
  }
}

BLCD_HD float sepfn_find_min(const SepFn& f, const DShape& A, const Sweep& sA, const DShape& B, const Sweep& sB, int* indexA, int* indexB, float t, bool a_static) {
  Xf xfA = a_static ? xf_identity() : sweep_xf(sA, t), xfB = sweep_xf(sB, t);
  if (f.type == 0) {
    V2 axisA = rmulT(xfA.q, f.axis);
    V2 axisB = rmulT(xfB.q, -f.axis);
    *indexA = support(A, axisA);
    *indexB = support(B, axisB);
    V2 pointA = xmul(xfA, A.v[*indexA]);
    V2 pointB = xmul(xfB, B.v[*indexB]);
    return dot(pointB - pointA, f.axis);
  } else if (f.type == 1) {
    V2 normal = rmul(xfA.q, f.axis);
    V2 pointA = xmul(xfA, f.localPoint);
    V2 axisB = rmulT(xfB.q, -normal);
    *indexA = -1;
    *indexB = support(B, axisB);
    V2 pointB = xmul(xfB, B.v[*indexB]);
    return dot(pointB - pointA, normal);
  } else {
    V2 normal = rmul(xfB.q, f.axis);
    V2 pointB = xmul(xfB, f.localPoint);
    V2 axisA = rmulT(xfA.q, -normal);
    *indexB = -1;
    *indexA = support(A, axisA);
    V2 pointA = xmul(xfA, A.v[*indexA]);
    return dot(pointA - pointB, normal);
  }
}

BLCD_HD float sepfn_eval(const SepFn& f, const DShape& A, const Sweep& sA, const DShape& B, const Sweep& sB, int indexA, int indexB, float t, bool a_static) {
  Xf xfA = a_static ? xf_identity() : sweep_xf(sA, t), xfB = sweep_xf(sB, t);
  if (f.type == 0) {
    V2 pointA = xmul(xfA, A.v[indexA]);
    V2 pointB = xmul(xfB, B.v[indexB]);
    return dot(pointB - pointA, f.axis);
  } else if (f.type == 1) {
    V2 normal = rmul(xfA.q, f.axis);
    V2 pointA = xmul(xfA, f.localPoint);
    V2 pointB = xmul(xfB, B.v[indexB]);
    return dot(pointB - pointA, normal);
  } else {
    V2 normal = rmul(xfB.q, f.axis);
    V2 pointB = xmul(xfB, f.localPoint);
    V2 pointA = xmul(xfA, A.v[indexA]);
    return dot(pointA - pointB, normal);
  }
}

enum { TOI_UNKNOWN = 0, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };

// b2TimeOfImpact with tMax = 1.  Returns the state and writes the fraction t.  a_static: fixture A belongs to a static
// body at the identity transform (a wall), whose sweep evaluates to the identity exactly at every t.
BLCD_HDN int time_of_impact(float* t_out, const DShape& A, Sweep sA, const DShape& B, Sweep sB, bool a_static) {
  const float tMax = 1.0f;
  int state = TOI_UNKNOWN;
  float t_res = tMax;
  sweep_normalize(sA);
  sweep_normalize(sB);
  float totalRadius = A.radius + B.radius;
  float target = fmaxb(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
  float tolerance = 0.25f * kLinearSlop;
  float t1 = 0.0f;
  int iter = 0;
  SimplexCache cache;
  cache.count = 0;
  cache.metric = 0.0f;
  for (;;) {
    Xf xfA = a_static ? xf_identity() : sweep_xf(sA, t1), xfB = sweep_xf(sB, t1);
    float distance = gjk_distance(cache, A, xfA, B, xfB);
    if (distance <= 0.0f) { state = TOI_OVERLAPPED; t_res = 0.0f; break; }
    if (distance < target + tolerance) { state = TOI_TOUCHING; t_res = t1; break; }
    SepFn fcn;
    sepfn_init(fcn, cache, A, sA, B, sB, t1, a_static);
    bool done = false;
    float t2 = tMax;
    int pushBackIter = 0;
    for (;;) {
      int indexA, indexB;
      float s2 = sepfn_find_min(fcn, A, sA, B, sB, &indexA, &indexB, t2, a_static);
      if (s2 > target + tolerance) { state = TOI_SEPARATED; t_res = tMax; done = true; break; }
      if (s2 > target - tolerance) { t1 = t2; break; }
      float s1 = sepfn_eval(fcn, A, sA, B, sB, indexA, indexB, t1, a_static);
      if (s1 < target - tolerance) { state = TOI_FAILED; t_res = t1; done = true; break; }
      if (s1 <= target + tolerance) { state = TOI_TOUCHING; t_res = t1; done = true; break; }
      int rootIterCount = 0;
      float a1 = t1, a2 = t2;
      for (;;) {
        float t;
        if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
        else t = 0.5f * (a1 + a2);
        ++rootIterCount;
        float s = sepfn_eval(fcn, A, sA, B, sB, indexA, indexB, t, a_static);
        if (absb(s - target) < tolerance) { t2 = t; break; }
        if (s > target) { a1 = t; s1 = s; }
        else { a2 = t; s2 = s; }
        if (rootIterCount == 50) break;
      }
      ++pushBackIter;
      if (pushBackIter == BLCD_MAX_VERTS) break;
    }
    ++iter;
    if (done) break;
    if (iter == 20) { state = TOI_FAILED; t_res = t1; break; }
  }
  *t_out = t_res;
  return state;
}

// tight AABB of a fixture under xf (b2Shape::ComputeAABB)
BLCD_HD Box shape_aabb(const DShape& s, const Xf& xf) {
  Box bb;
  if (s.type == SH_CIRCLE) {
    V2 p = xf.p + rmul(xf.q, s.v[0]);
    bb.lo = mk(p.x - s.radius, p.y - s.radius);
    bb.hi = mk(p.x + s.radius, p.y + s.radius);
    return bb;
  }
  V2 lo = xmul(xf, s.v[0]), hi = lo;
  for (int i = 1; i < s.count; ++i) {
    V2 v = xmul(xf, s.v[i]);
    lo = mk(fminb(lo.x, v.x), fminb(lo.y, v.y));
    hi = mk(fmaxb(hi.x, v.x), fmaxb(hi.y, v.y));
  }
  V2 r = mk(s.radius, s.radius);
  bb.lo = lo - r;
  bb.hi = hi + r;
  return bb;
}

}  // namespace BLCD_NS
