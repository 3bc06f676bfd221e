/*
 * b2_world.h -- TEST INFRASTRUCTURE (part of the CPU oracle), not product code.
 * `b2World::Step` for the subset of Box2D 2.3.x the reference exercises (boxLCD/world_env.py:446-452): contact manager
 * with fat-AABB pair bookkeeping, island DFS, b2Island::Solve / SolveTOI, b2ContactSolver, b2RevoluteJoint, sleeping and
 * continuous collision against static bodies.  Restated from upstream Box2D 2.3.x `Dynamics/b2World.cpp`,
 * `b2Island.cpp`, `b2ContactManager.cpp`, `Contacts/b2Contact.cpp`, `Contacts/b2ContactSolver.cpp`,
 * `Joints/b2RevoluteJoint.cpp`, `b2Body.cpp`, `b2Fixture.cpp`, `Collision/b2BroadPhase.cpp`, `b2DynamicTree.cpp`.
 * Parity: pybox2d cannot be run in this image, so there are no fp32 state vectors from it; the restatement is pinned
 * against what pybox2d DID produce -- the reference's recorded episodes (assets/envs/<Env>.gif), replayed frame by frame at
 * LCD resolution (tests/test_gif_episodes.py) and at the 8x colour resolution of the same frames (0.039 m per pixel,
 * tests/test_gif_hires.py), up to the oracle's own one-ulp chaos horizon.  See oracle/README.md.
 *
 * One fixture per body (all the reference ever creates).  The dynamic tree is replaced by a brute-force scan over fat
 * AABBs: the set of pairs it reports is the same, and Box2D sorts the pair buffer by proxy id before creating contacts,
 * so contact creation order is reproduced by using the fixture creation index as the proxy id (tree node ids of leaves
 * are monotonic in creation order for a world that never destroys a proxy; the reference builds a new b2World on every
 * reset, world_env.py:190-195).
 */
#pragma once
#include <algorithm>
#include <utility>
#include <vector>
#include "b2_collide.h"

namespace b2o {

enum BodyType { kStatic = 0, kDynamic = 2 };
enum LimitState { kInactiveLimit = 0, kAtLowerLimit = 1, kAtUpperLimit = 2, kEqualLimits = 3 };

struct WorldFlags {
  bool damping_2_3_0 = false, refface_2_3_0 = false, no_toi = false, no_sleep = false;
};

struct Counters {
  uint32_t contacts = 0, pos_iters = 0, toi_events = 0, toi_calls = 0, sleep_steps = 0, overflow = 0, manifold_points = 0, substeps = 0;
};

struct Body {
  int type = kDynamic;
  Transform xf;
  Sweep sweep;
  Vec2 v;
  float w = 0.0f;
  float mass = 0.0f, invMass = 0.0f, I = 0.0f, invI = 0.0f;
  float linearDamping = 0.0f, angularDamping = 0.0f;
  float sleepTime = 0.0f;
  bool awake = true, islandFlag = false;
  int islandIndex = 0;
  // fixture (exactly one)
  Shape shape;
  float friction = 0.2f, restitution = 0.0f, density = 0.0f;
  uint16_t categoryBits = 0x0001, maskBits = 0xFFFF;
  AABB fatAABB;
  std::vector<int> contactEdges;  // contact ids, newest first (b2Body::m_contactList)
  std::vector<int> jointEdges;    // joint ids, newest first (b2Body::m_jointList)

  void SynchronizeTransform() {
    xf.q.Set(sweep.a);
    xf.p = sweep.c - Mul(xf.q, sweep.localCenter);
  }
  void Advance(float alpha) {
    sweep.Advance(alpha);
    sweep.c = sweep.c0;
    sweep.a = sweep.a0;
    xf.q.Set(sweep.a);
    xf.p = sweep.c - Mul(xf.q, sweep.localCenter);
  }
  void SetAwake(bool flag) {
    if (flag) {
      if (!awake) { awake = true; sleepTime = 0.0f; }
    } else {
      awake = false; sleepTime = 0.0f;
      v = Vec2(0.0f, 0.0f); w = 0.0f;
    }
  }
};

struct Contact {
  int bodyA = 0, bodyB = 0;  // after the type-order swap of b2Contact::Create
  Manifold manifold;
  bool touching = false, enabled = true, islandFlag = false, toiFlag = false, alive = true;
  int toiCount = 0;
  float toi = 1.0f;
  float friction = 0.0f, restitution = 0.0f;
};

struct RevoluteJoint {
  int bodyA = 0, bodyB = 0;
  Vec2 localAnchorA, localAnchorB;
  float referenceAngle = 0.0f;
  bool enableLimit = false, enableMotor = false;
  float lowerAngle = 0.0f, upperAngle = 0.0f, maxMotorTorque = 0.0f, motorSpeed = 0.0f;
  Vec3 impulse;
  float motorImpulse = 0.0f;
  int limitState = kInactiveLimit;
  bool islandFlag = false;
  // solver temp
  int indexA = 0, indexB = 0;
  Vec2 rA, rB, localCenterA, localCenterB;
  float invMassA = 0, invMassB = 0, invIA = 0, invIB = 0;
  Mat33 mass;
  float motorMass = 0.0f;
};

struct Position { Vec2 c; float a; };
struct Velocity { Vec2 v; float w; };

struct TimeStep {
  float dt, inv_dt, dtRatio;
  int velocityIterations, positionIterations;
  bool warmStarting;
};

struct VelocityConstraintPoint {
  Vec2 rA, rB;
  float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias;
};

struct ContactVelocityConstraint {
  VelocityConstraintPoint points[2];
  Vec2 normal;
  Mat22 normalMass, K;
  int indexA, indexB;
  float invMassA, invMassB, invIA, invIB, friction, restitution;
  int pointCount, contactIndex;
};

struct ContactPositionConstraint {
  Vec2 localPoints[2];
  Vec2 localNormal, localPoint;
  int indexA, indexB;
  float invMassA, invMassB;
  Vec2 localCenterA, localCenterB;
  float invIA, invIB;
  int type;
  float radiusA, radiusB;
  int pointCount;
};

struct WorldManifold {
  Vec2 normal, points[2];
  void Initialize(const Manifold* manifold, const Transform& xfA, float radiusA, const Transform& xfB, float radiusB) {
    if (manifold->pointCount == 0) return;
    switch (manifold->type) {
      case kCircles: {
        normal = Vec2(1.0f, 0.0f);
        Vec2 pointA = Mul(xfA, manifold->localPoint);
        Vec2 pointB = Mul(xfB, manifold->points[0].localPoint);
        if (DistanceSquared(pointA, pointB) > kEpsilon * kEpsilon) {
          normal = pointB - pointA;
          normal.Normalize();
        }
        Vec2 cA = pointA + radiusA * normal;
        Vec2 cB = pointB - radiusB * normal;
        points[0] = 0.5f * (cA + cB);
      } break;
      case kFaceA: {
        normal = Mul(xfA.q, manifold->localNormal);
        Vec2 planePoint = Mul(xfA, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
          Vec2 clipPoint = Mul(xfB, manifold->points[i].localPoint);
          Vec2 cA = clipPoint + (radiusA - Dot(clipPoint - planePoint, normal)) * normal;
          Vec2 cB = clipPoint - radiusB * normal;
          points[i] = 0.5f * (cA + cB);
        }
      } break;
      case kFaceB: {
        normal = Mul(xfB.q, manifold->localNormal);
        Vec2 planePoint = Mul(xfB, manifold->localPoint);
        for (int i = 0; i < manifold->pointCount; ++i) {
          Vec2 clipPoint = Mul(xfA, manifold->points[i].localPoint);
          Vec2 cB = clipPoint + (radiusB - Dot(clipPoint - planePoint, normal)) * normal;
          Vec2 cA = clipPoint - radiusA * normal;
          points[i] = 0.5f * (cA + cB);
        }
        normal = -normal;
      } break;
    }
  }
};

struct PositionSolverManifold {
  Vec2 normal, point;
  float separation;
  void Initialize(const ContactPositionConstraint* pc, const Transform& xfA, const Transform& xfB, int index) {
    switch (pc->type) {
      case kCircles: {
        Vec2 pointA = Mul(xfA, pc->localPoint);
        Vec2 pointB = Mul(xfB, pc->localPoints[0]);
        normal = pointB - pointA;
        normal.Normalize();
        point = 0.5f * (pointA + pointB);
        separation = Dot(pointB - pointA, normal) - pc->radiusA - pc->radiusB;
      } break;
      case kFaceA: {
        normal = Mul(xfA.q, pc->localNormal);
        Vec2 planePoint = Mul(xfA, pc->localPoint);
        Vec2 clipPoint = Mul(xfB, pc->localPoints[index]);
        separation = Dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
        point = clipPoint;
      } break;
      default: {
        normal = Mul(xfB.q, pc->localNormal);
        Vec2 planePoint = Mul(xfB, pc->localPoint);
        Vec2 clipPoint = Mul(xfA, pc->localPoints[index]);
        separation = Dot(clipPoint - planePoint, normal) - pc->radiusA - pc->radiusB;
        point = clipPoint;
        normal = -normal;
      } break;
    }
  }
};

class World;

// b2ContactSolver
struct ContactSolver {
  TimeStep step;
  std::vector<Position>* positions;
  std::vector<Velocity>* velocities;
  std::vector<Contact*> contacts;
  std::vector<ContactVelocityConstraint> vcs;
  std::vector<ContactPositionConstraint> pcs;

  void Init(const TimeStep& st, const std::vector<int>& contactIds, World& world, std::vector<Position>* pos, std::vector<Velocity>* vel);
  void InitializeVelocityConstraints(World& world);
  void WarmStart();
  void SolveVelocityConstraints();
  void StoreImpulses();
  bool SolvePositionConstraints();
  bool SolveTOIPositionConstraints(int toiIndexA, int toiIndexB);
};

class World {
 public:
  std::vector<Body> bodies;        // creation order; body index == fixture index == proxy order
  std::vector<Contact> contacts;   // pool
  std::vector<int> contactList;    // ids, newest first (b2ContactManager::m_contactList)
  std::vector<RevoluteJoint> joints;
  std::vector<int> moveBuffer;
  Vec2 gravity{0.0f, -9.81f};
  float inv_dt0 = 0.0f;
  bool newFixture = false;
  WorldFlags flags;
  Counters counters;

  // ---- construction -------------------------------------------------------------------------------------------------
  int CreateBody(int type, Vec2 position, float angle, const Shape& shape, float density, float friction, float restitution,
                 uint16_t cat, uint16_t mask, float linDamp, float angDamp) {
    Body b;
    b.type = type;
    b.xf.p = position;
    b.xf.q.Set(angle);
    b.sweep.localCenter = Vec2(0.0f, 0.0f);
    b.sweep.c0 = b.sweep.c = position;
    b.sweep.a0 = b.sweep.a = angle;
    b.linearDamping = linDamp; b.angularDamping = angDamp;
    b.shape = shape; b.density = density; b.friction = friction; b.restitution = restitution;
    b.categoryBits = cat; b.maskBits = mask;
    // b2Body::CreateFixture -> CreateProxies (fat AABB = tight AABB +- b2_aabbExtension), ResetMassData
    AABB aabb = ComputeAABB(shape, b.xf);
    Vec2 r(kAabbExtension, kAabbExtension);
    b.fatAABB.lower = aabb.lower - r;
    b.fatAABB.upper = aabb.upper + r;
    if (type == kDynamic) {
      if (density > 0.0f) {
        MassData md = ComputeMass(shape, density);
        b.mass = md.mass;
        Vec2 localCenter = md.mass * md.center;
        b.I = md.I;
        if (b.mass > 0.0f) {
          b.invMass = 1.0f / b.mass;
          localCenter *= b.invMass;
        } else {
          b.mass = 1.0f; b.invMass = 1.0f;
        }
        if (b.I > 0.0f) {
          b.I -= b.mass * Dot(localCenter, localCenter);
          b.invI = 1.0f / b.I;
        } else {
          b.I = 0.0f; b.invI = 0.0f;
        }
        b.sweep.localCenter = localCenter;
        b.sweep.c0 = b.sweep.c = Mul(b.xf, localCenter);
      } else {
        b.mass = 1.0f; b.invMass = 1.0f;
      }
    }
    bodies.push_back(b);
    moveBuffer.push_back((int)bodies.size() - 1);
    newFixture = true;
    return (int)bodies.size() - 1;
  }

  int CreateRevoluteJoint(const RevoluteJoint& def) {
    joints.push_back(def);
    int id = (int)joints.size() - 1;
    bodies[def.bodyA].jointEdges.insert(bodies[def.bodyA].jointEdges.begin(), id);
    bodies[def.bodyB].jointEdges.insert(bodies[def.bodyB].jointEdges.begin(), id);
    return id;
  }

  // b2Body::SetTransform
  void SetTransform(int bi, Vec2 position, float angle) {
    Body& b = bodies[bi];
    b.xf.q.Set(angle);
    b.xf.p = position;
    b.sweep.c = Mul(b.xf, b.sweep.localCenter);
    b.sweep.a = angle;
    b.sweep.c0 = b.sweep.c;
    b.sweep.a0 = angle;
    SynchronizeFixture(bi, b.xf, b.xf);
  }

  void SetMotorSpeed(int ji, float speed) {
    RevoluteJoint& j = joints[ji];
    bodies[j.bodyA].SetAwake(true);
    bodies[j.bodyB].SetAwake(true);
    j.motorSpeed = speed;
  }

  // ---- broad phase --------------------------------------------------------------------------------------------------
  // b2Fixture::Synchronize + b2DynamicTree::MoveProxy
  void SynchronizeFixture(int bi, const Transform& xf1, const Transform& xf2) {
    Body& b = bodies[bi];
    AABB aabb1 = ComputeAABB(b.shape, xf1), aabb2 = ComputeAABB(b.shape, xf2), aabb;
    aabb.Combine(aabb1, aabb2);
    Vec2 displacement = xf2.p - xf1.p;
    if (b.fatAABB.Contains(aabb)) return;
    AABB fat = aabb;
    Vec2 r(kAabbExtension, kAabbExtension);
    fat.lower = fat.lower - r;
    fat.upper = fat.upper + r;
    Vec2 d = kAabbMultiplier * displacement;
    if (d.x < 0.0f) fat.lower.x += d.x; else fat.upper.x += d.x;
    if (d.y < 0.0f) fat.lower.y += d.y; else fat.upper.y += d.y;
    b.fatAABB = fat;
    moveBuffer.push_back(bi);
  }

  void SynchronizeFixtures(int bi) {
    Body& b = bodies[bi];
    Transform xf1;
    xf1.q.Set(b.sweep.a0);
    xf1.p = b.sweep.c0 - Mul(xf1.q, b.sweep.localCenter);
    SynchronizeFixture(bi, xf1, b.xf);
  }

  bool ShouldCollide(int a, int b) const {
    const Body &A = bodies[a], &B = bodies[b];
    if (A.type != kDynamic && B.type != kDynamic) return false;
    for (int ji : B.jointEdges) {  // collideConnected is false for every joint the reference creates
      const RevoluteJoint& j = joints[ji];
      if ((j.bodyA == a && j.bodyB == b) || (j.bodyA == b && j.bodyB == a)) return false;
    }
    return (A.maskBits & B.categoryBits) != 0 && (A.categoryBits & B.maskBits) != 0;
  }

  bool HasContact(int a, int b) const {
    for (int ci : bodies[b].contactEdges) {
      const Contact& c = contacts[ci];
      if ((c.bodyA == a && c.bodyB == b) || (c.bodyA == b && c.bodyB == a)) return true;
    }
    return false;
  }

  // b2BroadPhase::UpdatePairs + b2ContactManager::AddPair
  void FindNewContacts() {
    std::vector<std::pair<int, int>> pairs;
    for (int q : moveBuffer) {
      for (int o = 0; o < (int)bodies.size(); ++o) {
        if (o == q) continue;
        if (!TestOverlap(bodies[o].fatAABB, bodies[q].fatAABB)) continue;
        pairs.emplace_back(std::min(o, q), std::max(o, q));
      }
    }
    moveBuffer.clear();
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    for (auto& pr : pairs) {
      int a = pr.first, b = pr.second;
      if (HasContact(a, b)) continue;
      if (!ShouldCollide(b, a)) continue;
      // b2Contact::Create: s_registers[typeA][typeB].primary decides the order
      int ta = bodies[a].shape.type, tb = bodies[b].shape.type;
      bool primary = (ta == tb) || (ta == kPolygon && tb == kCircle) || (ta == kEdge && tb == kCircle) || (ta == kEdge && tb == kPolygon);
      Contact c;
      c.bodyA = primary ? a : b;
      c.bodyB = primary ? b : a;
      c.friction = sqrtf(bodies[a].friction * bodies[b].friction);
      c.restitution = Max(bodies[a].restitution, bodies[b].restitution);
      contacts.push_back(c);
      int id = (int)contacts.size() - 1;
      contactList.insert(contactList.begin(), id);
      bodies[c.bodyA].contactEdges.insert(bodies[c.bodyA].contactEdges.begin(), id);
      bodies[c.bodyB].contactEdges.insert(bodies[c.bodyB].contactEdges.begin(), id);
      bodies[c.bodyA].SetAwake(true);
      bodies[c.bodyB].SetAwake(true);
    }
  }

  void DestroyContact(int id) {
    Contact& c = contacts[id];
    if (c.manifold.pointCount > 0) {
      bodies[c.bodyA].SetAwake(true);
      bodies[c.bodyB].SetAwake(true);
    }
    auto rm = [id](std::vector<int>& v) { v.erase(std::remove(v.begin(), v.end(), id), v.end()); };
    rm(contactList);
    rm(bodies[c.bodyA].contactEdges);
    rm(bodies[c.bodyB].contactEdges);
    c.alive = false;
  }

  void Evaluate(Manifold* m, const Contact& c, const Transform& xfA, const Transform& xfB) const {
    const Shape &A = bodies[c.bodyA].shape, &B = bodies[c.bodyB].shape;
    if (A.type == kCircle && B.type == kCircle) CollideCircles(m, A, xfA, B, xfB);
    else if (A.type == kPolygon && B.type == kCircle) CollidePolygonAndCircle(m, A, xfA, B, xfB);
    else if (A.type == kPolygon && B.type == kPolygon) CollidePolygons(m, A, xfA, B, xfB, flags.refface_2_3_0);
    else if (A.type == kEdge && B.type == kCircle) CollideEdgeAndCircle(m, A, xfA, B, xfB);
    else if (A.type == kEdge && B.type == kPolygon) CollideEdgeAndPolygon(m, A, xfA, B, xfB);
    else m->pointCount = 0;
  }

  // b2Contact::Update
  void UpdateContact(int id) {
    Contact& c = contacts[id];
    Manifold oldManifold = c.manifold;
    c.enabled = true;
    bool wasTouching = c.touching;
    Body &bA = bodies[c.bodyA], &bB = bodies[c.bodyB];
    Evaluate(&c.manifold, c, bA.xf, bB.xf);
    bool touching = c.manifold.pointCount > 0;
    for (int i = 0; i < c.manifold.pointCount; ++i) {
      ManifoldPoint* mp2 = c.manifold.points + i;
      mp2->normalImpulse = 0.0f;
      mp2->tangentImpulse = 0.0f;
      for (int j = 0; j < oldManifold.pointCount; ++j) {
        const ManifoldPoint* mp1 = oldManifold.points + j;
        if (mp1->id.key() == mp2->id.key()) {
          mp2->normalImpulse = mp1->normalImpulse;
          mp2->tangentImpulse = mp1->tangentImpulse;
          break;
        }
      }
    }
    if (touching != wasTouching) {
      bA.SetAwake(true);
      bB.SetAwake(true);
    }
    c.touching = touching;
  }

  // b2ContactManager::Collide
  void Collide() {
    std::vector<int> list = contactList;
    for (int id : list) {
      Contact& c = contacts[id];
      const Body &bA = bodies[c.bodyA], &bB = bodies[c.bodyB];
      bool activeA = bA.awake && bA.type != kStatic;
      bool activeB = bB.awake && bB.type != kStatic;
      if (!activeA && !activeB) continue;
      if (!TestOverlap(bA.fatAABB, bB.fatAABB)) { DestroyContact(id); continue; }
      UpdateContact(id);
    }
  }

  // ---- joints -------------------------------------------------------------------------------------------------------
  void JointInitVelocityConstraints(RevoluteJoint& j, const TimeStep& step, std::vector<Position>& positions, std::vector<Velocity>& velocities) {
    const Body &bA = bodies[j.bodyA], &bB = bodies[j.bodyB];
    j.indexA = bA.islandIndex; j.indexB = bB.islandIndex;
    j.localCenterA = bA.sweep.localCenter; j.localCenterB = bB.sweep.localCenter;
    j.invMassA = bA.invMass; j.invMassB = bB.invMass;
    j.invIA = bA.invI; j.invIB = bB.invI;
    float aA = positions[j.indexA].a;
    Vec2 vA = velocities[j.indexA].v; float wA = velocities[j.indexA].w;
    float aB = positions[j.indexB].a;
    Vec2 vB = velocities[j.indexB].v; float wB = velocities[j.indexB].w;
    Rot qA(aA), qB(aB);
    j.rA = Mul(qA, j.localAnchorA - j.localCenterA);
    j.rB = Mul(qB, j.localAnchorB - j.localCenterB);
    float mA = j.invMassA, mB = j.invMassB, iA = j.invIA, iB = j.invIB;
    bool fixedRotation = (iA + iB == 0.0f);
    j.mass.ex.x = mA + mB + j.rA.y * j.rA.y * iA + j.rB.y * j.rB.y * iB;
    j.mass.ey.x = -j.rA.y * j.rA.x * iA - j.rB.y * j.rB.x * iB;
    j.mass.ez.x = -j.rA.y * iA - j.rB.y * iB;
    j.mass.ex.y = j.mass.ey.x;
    j.mass.ey.y = mA + mB + j.rA.x * j.rA.x * iA + j.rB.x * j.rB.x * iB;
    j.mass.ez.y = j.rA.x * iA + j.rB.x * iB;
    j.mass.ex.z = j.mass.ez.x;
    j.mass.ey.z = j.mass.ez.y;
    j.mass.ez.z = iA + iB;
    j.motorMass = iA + iB;
    if (j.motorMass > 0.0f) j.motorMass = 1.0f / j.motorMass;
    if (!j.enableMotor || fixedRotation) j.motorImpulse = 0.0f;
    if (j.enableLimit && !fixedRotation) {
      float jointAngle = aB - aA - j.referenceAngle;
      if (Abs(j.upperAngle - j.lowerAngle) < 2.0f * kAngularSlop) {
        j.limitState = kEqualLimits;
      } else if (jointAngle <= j.lowerAngle) {
        if (j.limitState != kAtLowerLimit) j.impulse.z = 0.0f;
        j.limitState = kAtLowerLimit;
      } else if (jointAngle >= j.upperAngle) {
        if (j.limitState != kAtUpperLimit) j.impulse.z = 0.0f;
        j.limitState = kAtUpperLimit;
      } else {
        j.limitState = kInactiveLimit;
        j.impulse.z = 0.0f;
      }
    } else {
      j.limitState = kInactiveLimit;
    }
    if (step.warmStarting) {
      j.impulse *= step.dtRatio;
      j.motorImpulse *= step.dtRatio;
      Vec2 P(j.impulse.x, j.impulse.y);
      vA -= mA * P;
      wA -= iA * (Cross(j.rA, P) + j.motorImpulse + j.impulse.z);
      vB += mB * P;
      wB += iB * (Cross(j.rB, P) + j.motorImpulse + j.impulse.z);
    } else {
      j.impulse = Vec3();
      j.motorImpulse = 0.0f;
    }
    velocities[j.indexA].v = vA; velocities[j.indexA].w = wA;
    velocities[j.indexB].v = vB; velocities[j.indexB].w = wB;
  }

  void JointSolveVelocityConstraints(RevoluteJoint& j, const TimeStep& step, std::vector<Velocity>& velocities) {
    Vec2 vA = velocities[j.indexA].v; float wA = velocities[j.indexA].w;
    Vec2 vB = velocities[j.indexB].v; float wB = velocities[j.indexB].w;
    float mA = j.invMassA, mB = j.invMassB, iA = j.invIA, iB = j.invIB;
    bool fixedRotation = (iA + iB == 0.0f);
    if (j.enableMotor && j.limitState != kEqualLimits && !fixedRotation) {
      float Cdot = wB - wA - j.motorSpeed;
      float impulse = -j.motorMass * Cdot;
      float oldImpulse = j.motorImpulse;
      float maxImpulse = step.dt * j.maxMotorTorque;
      j.motorImpulse = Clamp(oldImpulse + impulse, -maxImpulse, maxImpulse);
      impulse = j.motorImpulse - oldImpulse;
      wA -= iA * impulse;
      wB += iB * impulse;
    }
    if (j.enableLimit && j.limitState != kInactiveLimit && !fixedRotation) {
      Vec2 Cdot1 = vB + Cross(wB, j.rB) - vA - Cross(wA, j.rA);
      float Cdot2 = wB - wA;
      Vec3 Cdot(Cdot1.x, Cdot1.y, Cdot2);
      Vec3 impulse = -j.mass.Solve33(Cdot);
      if (j.limitState == kEqualLimits) {
        j.impulse += impulse;
      } else if (j.limitState == kAtLowerLimit) {
        float newImpulse = j.impulse.z + impulse.z;
        if (newImpulse < 0.0f) {
          Vec2 rhs = -Cdot1 + j.impulse.z * Vec2(j.mass.ez.x, j.mass.ez.y);
          Vec2 reduced = j.mass.Solve22(rhs);
          impulse.x = reduced.x; impulse.y = reduced.y; impulse.z = -j.impulse.z;
          j.impulse.x += reduced.x; j.impulse.y += reduced.y; j.impulse.z = 0.0f;
        } else {
          j.impulse += impulse;
        }
      } else if (j.limitState == kAtUpperLimit) {
        float newImpulse = j.impulse.z + impulse.z;
        if (newImpulse > 0.0f) {
          Vec2 rhs = -Cdot1 + j.impulse.z * Vec2(j.mass.ez.x, j.mass.ez.y);
          Vec2 reduced = j.mass.Solve22(rhs);
          impulse.x = reduced.x; impulse.y = reduced.y; impulse.z = -j.impulse.z;
          j.impulse.x += reduced.x; j.impulse.y += reduced.y; j.impulse.z = 0.0f;
        } else {
          j.impulse += impulse;
        }
      }
      Vec2 P(impulse.x, impulse.y);
      vA -= mA * P;
      wA -= iA * (Cross(j.rA, P) + impulse.z);
      vB += mB * P;
      wB += iB * (Cross(j.rB, P) + impulse.z);
    } else {
      Vec2 Cdot = vB + Cross(wB, j.rB) - vA - Cross(wA, j.rA);
      Vec2 impulse = j.mass.Solve22(-Cdot);
      j.impulse.x += impulse.x;
      j.impulse.y += impulse.y;
      vA -= mA * impulse;
      wA -= iA * Cross(j.rA, impulse);
      vB += mB * impulse;
      wB += iB * Cross(j.rB, impulse);
    }
    velocities[j.indexA].v = vA; velocities[j.indexA].w = wA;
    velocities[j.indexB].v = vB; velocities[j.indexB].w = wB;
  }

  bool JointSolvePositionConstraints(RevoluteJoint& j, std::vector<Position>& positions) {
    Vec2 cA = positions[j.indexA].c; float aA = positions[j.indexA].a;
    Vec2 cB = positions[j.indexB].c; float aB = positions[j.indexB].a;
    Rot qA(aA), qB(aB);
    float angularError = 0.0f, positionError = 0.0f;
    bool fixedRotation = (j.invIA + j.invIB == 0.0f);
    if (j.enableLimit && j.limitState != kInactiveLimit && !fixedRotation) {
      float angle = aB - aA - j.referenceAngle;
      float limitImpulse = 0.0f;
      if (j.limitState == kEqualLimits) {
        float C = Clamp(angle - j.lowerAngle, -kMaxAngularCorrection, kMaxAngularCorrection);
        limitImpulse = -j.motorMass * C;
        angularError = Abs(C);
      } else if (j.limitState == kAtLowerLimit) {
        float C = angle - j.lowerAngle;
        angularError = -C;
        C = Clamp(C + kAngularSlop, -kMaxAngularCorrection, 0.0f);
        limitImpulse = -j.motorMass * C;
      } else if (j.limitState == kAtUpperLimit) {
        float C = angle - j.upperAngle;
        angularError = C;
        C = Clamp(C - kAngularSlop, 0.0f, kMaxAngularCorrection);
        limitImpulse = -j.motorMass * C;
      }
      aA -= j.invIA * limitImpulse;
      aB += j.invIB * limitImpulse;
    }
    {
      qA.Set(aA);
      qB.Set(aB);
      Vec2 rA = Mul(qA, j.localAnchorA - j.localCenterA);
      Vec2 rB = Mul(qB, j.localAnchorB - j.localCenterB);
      Vec2 C = cB + rB - cA - rA;
      positionError = C.Length();
      float mA = j.invMassA, mB = j.invMassB, iA = j.invIA, iB = j.invIB;
      Mat22 K;
      K.ex.x = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
      K.ex.y = -iA * rA.x * rA.y - iB * rB.x * rB.y;
      K.ey.x = K.ex.y;
      K.ey.y = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
      Vec2 impulse = -K.Solve(C);
      cA -= mA * impulse;
      aA -= iA * Cross(rA, impulse);
      cB += mB * impulse;
      aB += iB * Cross(rB, impulse);
    }
    positions[j.indexA].c = cA; positions[j.indexA].a = aA;
    positions[j.indexB].c = cB; positions[j.indexB].a = aB;
    return positionError <= kLinearSlop && angularError <= kAngularSlop;
  }

  // ---- islands ------------------------------------------------------------------------------------------------------
  void IslandSolve(const std::vector<int>& ibodies, const std::vector<int>& icontacts, const std::vector<int>& ijoints, const TimeStep& step) {
    float h = step.dt;
    std::vector<Position> positions(ibodies.size());
    std::vector<Velocity> velocities(ibodies.size());
    for (size_t i = 0; i < ibodies.size(); ++i) {
      Body& b = bodies[ibodies[i]];
      Vec2 c = b.sweep.c; float a = b.sweep.a;
      Vec2 v = b.v; float w = b.w;
      b.sweep.c0 = b.sweep.c;
      b.sweep.a0 = b.sweep.a;
      if (b.type == kDynamic) {
        // gravityScale = 1, no applied force / torque on this path
        v += h * (1.0f * gravity + b.invMass * Vec2(0.0f, 0.0f));
        w += h * b.invI * 0.0f;
        if (flags.damping_2_3_0) {
          v *= Clamp(1.0f - h * b.linearDamping, 0.0f, 1.0f);
          w *= Clamp(1.0f - h * b.angularDamping, 0.0f, 1.0f);
        } else {
          v *= 1.0f / (1.0f + h * b.linearDamping);
          w *= 1.0f / (1.0f + h * b.angularDamping);
        }
      }
      positions[i].c = c; positions[i].a = a;
      velocities[i].v = v; velocities[i].w = w;
    }
    ContactSolver cs;
    cs.Init(step, icontacts, *this, &positions, &velocities);
    cs.InitializeVelocityConstraints(*this);
    if (step.warmStarting) cs.WarmStart();
    for (int ji : ijoints) JointInitVelocityConstraints(joints[ji], step, positions, velocities);
    for (int i = 0; i < step.velocityIterations; ++i) {
      for (int ji : ijoints) JointSolveVelocityConstraints(joints[ji], step, velocities);
      cs.SolveVelocityConstraints();
    }
    cs.StoreImpulses();
    for (size_t i = 0; i < ibodies.size(); ++i) {
      Vec2 c = positions[i].c; float a = positions[i].a;
      Vec2 v = velocities[i].v; float w = velocities[i].w;
      Vec2 translation = h * v;
      if (Dot(translation, translation) > kMaxTranslationSquared) {
        float ratio = kMaxTranslation / translation.Length();
        v *= ratio;
      }
      float rotation = h * w;
      if (rotation * rotation > kMaxRotationSquared) {
        float ratio = kMaxRotation / Abs(rotation);
        w *= ratio;
      }
      c += h * v;
      a += h * w;
      positions[i].c = c; positions[i].a = a;
      velocities[i].v = v; velocities[i].w = w;
    }
    bool positionSolved = false;
    for (int i = 0; i < step.positionIterations; ++i) {
      ++counters.pos_iters;
      bool contactsOkay = cs.SolvePositionConstraints();
      bool jointsOkay = true;
      for (int ji : ijoints) {
        bool jointOkay = JointSolvePositionConstraints(joints[ji], positions);
        jointsOkay = jointsOkay && jointOkay;
      }
      if (contactsOkay && jointsOkay) { positionSolved = true; break; }
    }
    for (size_t i = 0; i < ibodies.size(); ++i) {
      Body& b = bodies[ibodies[i]];
      b.sweep.c = positions[i].c;
      b.sweep.a = positions[i].a;
      b.v = velocities[i].v;
      b.w = velocities[i].w;
      b.SynchronizeTransform();
    }
    if (!flags.no_sleep) {
      float minSleepTime = kMaxFloat;
      const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
      const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
      for (int bi : ibodies) {
        Body& b = bodies[bi];
        if (b.type == kStatic) continue;
        if (b.w * b.w > angTolSqr || Dot(b.v, b.v) > linTolSqr) {
          b.sleepTime = 0.0f;
          minSleepTime = 0.0f;
        } else {
          b.sleepTime += h;
          minSleepTime = Min(minSleepTime, b.sleepTime);
        }
      }
      if (minSleepTime >= kTimeToSleep && positionSolved) {
        for (int bi : ibodies) bodies[bi].SetAwake(false);
      }
    }
  }

  void IslandSolveTOI(const std::vector<int>& ibodies, const std::vector<int>& icontacts, const TimeStep& subStep, int toiIndexA, int toiIndexB) {
    std::vector<Position> positions(ibodies.size());
    std::vector<Velocity> velocities(ibodies.size());
    for (size_t i = 0; i < ibodies.size(); ++i) {
      const Body& b = bodies[ibodies[i]];
      positions[i].c = b.sweep.c; positions[i].a = b.sweep.a;
      velocities[i].v = b.v; velocities[i].w = b.w;
    }
    ContactSolver cs;
    cs.Init(subStep, icontacts, *this, &positions, &velocities);
    for (int i = 0; i < subStep.positionIterations; ++i) {
      bool contactsOkay = cs.SolveTOIPositionConstraints(toiIndexA, toiIndexB);
      if (contactsOkay) break;
    }
    bodies[ibodies[toiIndexA]].sweep.c0 = positions[toiIndexA].c;
    bodies[ibodies[toiIndexA]].sweep.a0 = positions[toiIndexA].a;
    bodies[ibodies[toiIndexB]].sweep.c0 = positions[toiIndexB].c;
    bodies[ibodies[toiIndexB]].sweep.a0 = positions[toiIndexB].a;
    cs.InitializeVelocityConstraints(*this);
    for (int i = 0; i < subStep.velocityIterations; ++i) cs.SolveVelocityConstraints();
    float h = subStep.dt;
    for (size_t i = 0; i < ibodies.size(); ++i) {
      Vec2 c = positions[i].c; float a = positions[i].a;
      Vec2 v = velocities[i].v; float w = velocities[i].w;
      Vec2 translation = h * v;
      if (Dot(translation, translation) > kMaxTranslationSquared) {
        float ratio = kMaxTranslation / translation.Length();
        v *= ratio;
      }
      float rotation = h * w;
      if (rotation * rotation > kMaxRotationSquared) {
        float ratio = kMaxRotation / Abs(rotation);
        w *= ratio;
      }
      c += h * v;
      a += h * w;
      Body& b = bodies[ibodies[i]];
      b.sweep.c = c; b.sweep.a = a;
      b.v = v; b.w = w;
      b.SynchronizeTransform();
    }
  }

  // b2World::Solve
  void Solve(const TimeStep& step) {
    for (Body& b : bodies) b.islandFlag = false;
    for (int id : contactList) contacts[id].islandFlag = false;
    for (RevoluteJoint& j : joints) j.islandFlag = false;
    std::vector<int> stack;
    for (int seed = (int)bodies.size() - 1; seed >= 0; --seed) {  // m_bodyList is newest first
      Body& sb = bodies[seed];
      if (sb.islandFlag) continue;
      if (!sb.awake) continue;
      if (sb.type == kStatic) continue;
      std::vector<int> ibodies, icontacts, ijoints;
      stack.clear();
      stack.push_back(seed);
      sb.islandFlag = true;
      while (!stack.empty()) {
        int bi = stack.back();
        stack.pop_back();
        Body& b = bodies[bi];
        b.islandIndex = (int)ibodies.size();
        ibodies.push_back(bi);
        b.SetAwake(true);
        if (b.type == kStatic) continue;
        for (int ci : b.contactEdges) {
          Contact& c = contacts[ci];
          if (c.islandFlag) continue;
          if (!c.enabled || !c.touching) continue;
          icontacts.push_back(ci);
          c.islandFlag = true;
          int other = c.bodyA == bi ? c.bodyB : c.bodyA;
          if (bodies[other].islandFlag) continue;
          stack.push_back(other);
          bodies[other].islandFlag = true;
        }
        for (int ji : b.jointEdges) {
          RevoluteJoint& j = joints[ji];
          if (j.islandFlag) continue;
          int other = j.bodyA == bi ? j.bodyB : j.bodyA;
          ijoints.push_back(ji);
          j.islandFlag = true;
          if (bodies[other].islandFlag) continue;
          stack.push_back(other);
          bodies[other].islandFlag = true;
        }
      }
      counters.contacts += (uint32_t)icontacts.size();
      IslandSolve(ibodies, icontacts, ijoints, step);
      for (int bi : ibodies) {
        if (bodies[bi].type == kStatic) bodies[bi].islandFlag = false;
      }
    }
    for (int bi = (int)bodies.size() - 1; bi >= 0; --bi) {
      Body& b = bodies[bi];
      if (!b.islandFlag) continue;
      if (b.type == kStatic) continue;
      SynchronizeFixtures(bi);
    }
    FindNewContacts();
  }

  // b2World::SolveTOI (stepComplete is always true here: sub-stepping is off)
  void SolveTOI(const TimeStep& step) {
    for (Body& b : bodies) { b.islandFlag = false; b.sweep.alpha0 = 0.0f; }
    for (int id : contactList) {
      Contact& c = contacts[id];
      c.toiFlag = false; c.islandFlag = false;
      c.toiCount = 0;
      c.toi = 1.0f;
    }
    for (;;) {
      int minContact = -1;
      float minAlpha = 1.0f;
      for (int id : contactList) {
        Contact& c = contacts[id];
        if (!c.enabled) continue;
        if (c.toiCount > kMaxSubSteps) continue;
        float alpha = 1.0f;
        if (c.toiFlag) {
          alpha = c.toi;
        } else {
          Body &bA = bodies[c.bodyA], &bB = bodies[c.bodyB];
          bool activeA = bA.awake && bA.type != kStatic;
          bool activeB = bB.awake && bB.type != kStatic;
          if (!activeA && !activeB) continue;
          bool collideA = bA.type != kDynamic;  // no bullets on this path
          bool collideB = bB.type != kDynamic;
          if (!collideA && !collideB) continue;
          float alpha0 = bA.sweep.alpha0;
          if (bA.sweep.alpha0 < bB.sweep.alpha0) {
            alpha0 = bB.sweep.alpha0;
            bA.sweep.Advance(alpha0);
          } else if (bB.sweep.alpha0 < bA.sweep.alpha0) {
            alpha0 = bA.sweep.alpha0;
            bB.sweep.Advance(alpha0);
          }
          DistanceProxy proxyA, proxyB;
          proxyA.Set(bA.shape);
          proxyB.Set(bB.shape);
          ToiOutput output;
          ++counters.toi_calls;
          TimeOfImpact(&output, &proxyA, &proxyB, bA.sweep, bB.sweep, 1.0f);
          float beta = output.t;
          if (output.state == kToiTouching) alpha = Min(alpha0 + (1.0f - alpha0) * beta, 1.0f);
          else alpha = 1.0f;
          c.toi = alpha;
          c.toiFlag = true;
        }
        if (alpha < minAlpha) { minContact = id; minAlpha = alpha; }
      }
      if (minContact < 0 || 1.0f - 10.0f * kEpsilon < minAlpha) break;
      Contact& mc = contacts[minContact];
      int iA = mc.bodyA, iB = mc.bodyB;
      Body &bA = bodies[iA], &bB = bodies[iB];
      Sweep backup1 = bA.sweep, backup2 = bB.sweep;
      bA.Advance(minAlpha);
      bB.Advance(minAlpha);
      UpdateContact(minContact);
      mc.toiFlag = false;
      ++mc.toiCount;
      if (!mc.enabled || !mc.touching) {
        mc.enabled = false;
        bA.sweep = backup1;
        bB.sweep = backup2;
        bA.SynchronizeTransform();
        bB.SynchronizeTransform();
        continue;
      }
      ++counters.toi_events;
      bA.SetAwake(true);
      bB.SetAwake(true);
      std::vector<int> ibodies, icontacts;
      bA.islandIndex = 0; ibodies.push_back(iA);
      bB.islandIndex = 1; ibodies.push_back(iB);
      icontacts.push_back(minContact);
      bA.islandFlag = true; bB.islandFlag = true; mc.islandFlag = true;
      int two[2] = {iA, iB};
      for (int k = 0; k < 2; ++k) {
        int bi = two[k];
        Body& body = bodies[bi];
        if (body.type != kDynamic) continue;
        for (int ci : body.contactEdges) {
          if ((int)ibodies.size() == 2 * kMaxTOIContacts) break;
          if ((int)icontacts.size() == kMaxTOIContacts) break;
          Contact& c = contacts[ci];
          if (c.islandFlag) continue;
          int oi = c.bodyA == bi ? c.bodyB : c.bodyA;
          Body& other = bodies[oi];
          if (other.type == kDynamic) continue;  // neither side is a bullet
          Sweep backup = other.sweep;
          if (!other.islandFlag) other.Advance(minAlpha);
          UpdateContact(ci);
          if (!c.enabled || !c.touching) {
            other.sweep = backup;
            other.SynchronizeTransform();
            continue;
          }
          c.islandFlag = true;
          icontacts.push_back(ci);
          if (other.islandFlag) continue;
          other.islandFlag = true;
          if (other.type != kStatic) other.SetAwake(true);
          other.islandIndex = (int)ibodies.size();
          ibodies.push_back(oi);
        }
      }
      TimeStep subStep;
      subStep.dt = (1.0f - minAlpha) * step.dt;
      subStep.inv_dt = 1.0f / subStep.dt;
      subStep.dtRatio = 1.0f;
      subStep.positionIterations = 20;
      subStep.velocityIterations = step.velocityIterations;
      subStep.warmStarting = false;
      IslandSolveTOI(ibodies, icontacts, subStep, bodies[iA].islandIndex, bodies[iB].islandIndex);
      for (int bi : ibodies) {
        Body& body = bodies[bi];
        body.islandFlag = false;
        if (body.type != kDynamic) continue;
        SynchronizeFixtures(bi);
        for (int ci : body.contactEdges) {
          contacts[ci].toiFlag = false;
          contacts[ci].islandFlag = false;
        }
      }
      FindNewContacts();
    }
  }

  // b2World::Step
  void Step(float dt, int velocityIterations, int positionIterations) {
    if (newFixture) {
      FindNewContacts();
      newFixture = false;
    }
    TimeStep step;
    step.dt = dt;
    step.velocityIterations = velocityIterations;
    step.positionIterations = positionIterations;
    step.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
    step.dtRatio = inv_dt0 * dt;
    step.warmStarting = true;
    Collide();
    if (step.dt > 0.0f) Solve(step);
    if (!flags.no_toi && step.dt > 0.0f) SolveTOI(step);
    if (step.dt > 0.0f) inv_dt0 = step.inv_dt;
    ++counters.substeps;
    for (int id : contactList) counters.manifold_points += (uint32_t)contacts[id].manifold.pointCount;
    for (const Body& b : bodies) if (b.type == kDynamic && !b.awake) ++counters.sleep_steps;
  }
};

// ---- b2ContactSolver -------------------------------------------------------------------------------------------------
inline void ContactSolver::Init(const TimeStep& st, const std::vector<int>& contactIds, World& world, std::vector<Position>* pos, std::vector<Velocity>* vel) {
  step = st;
  positions = pos;
  velocities = vel;
  int count = (int)contactIds.size();
  contacts.resize(count);
  vcs.resize(count);
  pcs.resize(count);
  for (int i = 0; i < count; ++i) {
    Contact* contact = &world.contacts[contactIds[i]];
    contacts[i] = contact;
    const Body &bodyA = world.bodies[contact->bodyA], &bodyB = world.bodies[contact->bodyB];
    const Manifold* manifold = &contact->manifold;
    int pointCount = manifold->pointCount;
    ContactVelocityConstraint* vc = &vcs[i];
    vc->friction = contact->friction;
    vc->restitution = contact->restitution;
    vc->indexA = bodyA.islandIndex; vc->indexB = bodyB.islandIndex;
    vc->invMassA = bodyA.invMass; vc->invMassB = bodyB.invMass;
    vc->invIA = bodyA.invI; vc->invIB = bodyB.invI;
    vc->contactIndex = i;
    vc->pointCount = pointCount;
    vc->K = Mat22(); vc->normalMass = Mat22();
    ContactPositionConstraint* pc = &pcs[i];
    pc->indexA = bodyA.islandIndex; pc->indexB = bodyB.islandIndex;
    pc->invMassA = bodyA.invMass; pc->invMassB = bodyB.invMass;
    pc->localCenterA = bodyA.sweep.localCenter; pc->localCenterB = bodyB.sweep.localCenter;
    pc->invIA = bodyA.invI; pc->invIB = bodyB.invI;
    pc->localNormal = manifold->localNormal;
    pc->localPoint = manifold->localPoint;
    pc->pointCount = pointCount;
    pc->radiusA = bodyA.shape.radius; pc->radiusB = bodyB.shape.radius;
    pc->type = manifold->type;
    for (int j = 0; j < pointCount; ++j) {
      const ManifoldPoint* cp = manifold->points + j;
      VelocityConstraintPoint* vcp = vc->points + j;
      if (step.warmStarting) {
        vcp->normalImpulse = step.dtRatio * cp->normalImpulse;
        vcp->tangentImpulse = step.dtRatio * cp->tangentImpulse;
      } else {
        vcp->normalImpulse = 0.0f;
        vcp->tangentImpulse = 0.0f;
      }
      vcp->rA = Vec2(); vcp->rB = Vec2();
      vcp->normalMass = 0.0f; vcp->tangentMass = 0.0f; vcp->velocityBias = 0.0f;
      pc->localPoints[j] = cp->localPoint;
    }
  }
}

inline void ContactSolver::InitializeVelocityConstraints(World&) {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    ContactPositionConstraint* pc = &pcs[i];
    float radiusA = pc->radiusA, radiusB = pc->radiusB;
    const Manifold* manifold = &contacts[vc->contactIndex]->manifold;
    int indexA = vc->indexA, indexB = vc->indexB;
    float mA = vc->invMassA, mB = vc->invMassB, iA = vc->invIA, iB = vc->invIB;
    Vec2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
    Vec2 cA = (*positions)[indexA].c; float aA = (*positions)[indexA].a;
    Vec2 vA = (*velocities)[indexA].v; float wA = (*velocities)[indexA].w;
    Vec2 cB = (*positions)[indexB].c; float aB = (*positions)[indexB].a;
    Vec2 vB = (*velocities)[indexB].v; float wB = (*velocities)[indexB].w;
    Transform xfA, xfB;
    xfA.q.Set(aA);
    xfB.q.Set(aB);
    xfA.p = cA - Mul(xfA.q, localCenterA);
    xfB.p = cB - Mul(xfB.q, localCenterB);
    WorldManifold worldManifold;
    worldManifold.Initialize(manifold, xfA, radiusA, xfB, radiusB);
    vc->normal = worldManifold.normal;
    int pointCount = vc->pointCount;
    for (int j = 0; j < pointCount; ++j) {
      VelocityConstraintPoint* vcp = vc->points + j;
      vcp->rA = worldManifold.points[j] - cA;
      vcp->rB = worldManifold.points[j] - cB;
      float rnA = Cross(vcp->rA, vc->normal), rnB = Cross(vcp->rB, vc->normal);
      float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      vcp->normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
      Vec2 tangent = Cross(vc->normal, 1.0f);
      float rtA = Cross(vcp->rA, tangent), rtB = Cross(vcp->rB, tangent);
      float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
      vcp->tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
      vcp->velocityBias = 0.0f;
      float vRel = Dot(vc->normal, vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA));
      if (vRel < -kVelocityThreshold) vcp->velocityBias = -vc->restitution * vRel;
    }
    if (vc->pointCount == 2) {
      VelocityConstraintPoint* vcp1 = vc->points + 0;
      VelocityConstraintPoint* vcp2 = vc->points + 1;
      float rn1A = Cross(vcp1->rA, vc->normal), rn1B = Cross(vcp1->rB, vc->normal);
      float rn2A = Cross(vcp2->rA, vc->normal), rn2B = Cross(vcp2->rB, vc->normal);
      float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
      float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
      float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
      const float k_maxConditionNumber = 1000.0f;
      if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
        vc->K.ex = Vec2(k11, k12);
        vc->K.ey = Vec2(k12, k22);
        vc->normalMass = vc->K.GetInverse();
      } else {
        vc->pointCount = 1;
      }
    }
  }
}

inline void ContactSolver::WarmStart() {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    int indexA = vc->indexA, indexB = vc->indexB;
    float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
    Vec2 vA = (*velocities)[indexA].v; float wA = (*velocities)[indexA].w;
    Vec2 vB = (*velocities)[indexB].v; float wB = (*velocities)[indexB].w;
    Vec2 normal = vc->normal;
    Vec2 tangent = Cross(normal, 1.0f);
    for (int j = 0; j < vc->pointCount; ++j) {
      VelocityConstraintPoint* vcp = vc->points + j;
      Vec2 P = vcp->normalImpulse * normal + vcp->tangentImpulse * tangent;
      wA -= iA * Cross(vcp->rA, P);
      vA -= mA * P;
      wB += iB * Cross(vcp->rB, P);
      vB += mB * P;
    }
    (*velocities)[indexA].v = vA; (*velocities)[indexA].w = wA;
    (*velocities)[indexB].v = vB; (*velocities)[indexB].w = wB;
  }
}

inline void ContactSolver::SolveVelocityConstraints() {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    int indexA = vc->indexA, indexB = vc->indexB;
    float mA = vc->invMassA, iA = vc->invIA, mB = vc->invMassB, iB = vc->invIB;
    int pointCount = vc->pointCount;
    Vec2 vA = (*velocities)[indexA].v; float wA = (*velocities)[indexA].w;
    Vec2 vB = (*velocities)[indexB].v; float wB = (*velocities)[indexB].w;
    Vec2 normal = vc->normal;
    Vec2 tangent = Cross(normal, 1.0f);
    float friction = vc->friction;
    for (int j = 0; j < pointCount; ++j) {
      VelocityConstraintPoint* vcp = vc->points + j;
      Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
      float vt = Dot(dv, tangent) - 0.0f;  // tangentSpeed = 0
      float lambda = vcp->tangentMass * (-vt);
      float maxFriction = friction * vcp->normalImpulse;
      float newImpulse = Clamp(vcp->tangentImpulse + lambda, -maxFriction, maxFriction);
      lambda = newImpulse - vcp->tangentImpulse;
      vcp->tangentImpulse = newImpulse;
      Vec2 P = lambda * tangent;
      vA -= mA * P;
      wA -= iA * Cross(vcp->rA, P);
      vB += mB * P;
      wB += iB * Cross(vcp->rB, P);
    }
    if (vc->pointCount == 1) {
      VelocityConstraintPoint* vcp = vc->points + 0;
      Vec2 dv = vB + Cross(wB, vcp->rB) - vA - Cross(wA, vcp->rA);
      float vn = Dot(dv, normal);
      float lambda = -vcp->normalMass * (vn - vcp->velocityBias);
      float newImpulse = Max(vcp->normalImpulse + lambda, 0.0f);
      lambda = newImpulse - vcp->normalImpulse;
      vcp->normalImpulse = newImpulse;
      Vec2 P = lambda * normal;
      vA -= mA * P;
      wA -= iA * Cross(vcp->rA, P);
      vB += mB * P;
      wB += iB * Cross(vcp->rB, P);
    } else {
      VelocityConstraintPoint* cp1 = vc->points + 0;
      VelocityConstraintPoint* cp2 = vc->points + 1;
      Vec2 a(cp1->normalImpulse, cp2->normalImpulse);
      Vec2 dv1 = vB + Cross(wB, cp1->rB) - vA - Cross(wA, cp1->rA);
      Vec2 dv2 = vB + Cross(wB, cp2->rB) - vA - Cross(wA, cp2->rA);
      float vn1 = Dot(dv1, normal);
      float vn2 = Dot(dv2, normal);
      Vec2 b;
      b.x = vn1 - cp1->velocityBias;
      b.y = vn2 - cp2->velocityBias;
      b -= Mul(vc->K, a);
      auto apply = [&](const Vec2& x) {
        Vec2 d = x - a;
        Vec2 P1 = d.x * normal;
        Vec2 P2 = d.y * normal;
        vA -= mA * (P1 + P2);
        wA -= iA * (Cross(cp1->rA, P1) + Cross(cp2->rA, P2));
        vB += mB * (P1 + P2);
        wB += iB * (Cross(cp1->rB, P1) + Cross(cp2->rB, P2));
        cp1->normalImpulse = x.x;
        cp2->normalImpulse = x.y;
      };
      for (;;) {
        Vec2 x = -Mul(vc->normalMass, b);
        if (x.x >= 0.0f && x.y >= 0.0f) { apply(x); break; }
        x.x = -cp1->normalMass * b.x;
        x.y = 0.0f;
        vn1 = 0.0f;
        vn2 = vc->K.ex.y * x.x + b.y;
        if (x.x >= 0.0f && vn2 >= 0.0f) { apply(x); break; }
        x.x = 0.0f;
        x.y = -cp2->normalMass * b.y;
        vn1 = vc->K.ey.x * x.y + b.x;
        vn2 = 0.0f;
        if (x.y >= 0.0f && vn1 >= 0.0f) { apply(x); break; }
        x.x = 0.0f;
        x.y = 0.0f;
        vn1 = b.x;
        vn2 = b.y;
        if (vn1 >= 0.0f && vn2 >= 0.0f) { apply(x); break; }
        break;
      }
    }
    (*velocities)[indexA].v = vA; (*velocities)[indexA].w = wA;
    (*velocities)[indexB].v = vB; (*velocities)[indexB].w = wB;
  }
}

inline void ContactSolver::StoreImpulses() {
  for (size_t i = 0; i < vcs.size(); ++i) {
    ContactVelocityConstraint* vc = &vcs[i];
    Manifold* manifold = &contacts[vc->contactIndex]->manifold;
    for (int j = 0; j < vc->pointCount; ++j) {
      manifold->points[j].normalImpulse = vc->points[j].normalImpulse;
      manifold->points[j].tangentImpulse = vc->points[j].tangentImpulse;
    }
  }
}

inline bool ContactSolver::SolvePositionConstraints() {
  float minSeparation = 0.0f;
  for (size_t i = 0; i < pcs.size(); ++i) {
    ContactPositionConstraint* pc = &pcs[i];
    int indexA = pc->indexA, indexB = pc->indexB;
    Vec2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
    float mA = pc->invMassA, iA = pc->invIA, mB = pc->invMassB, iB = pc->invIB;
    int pointCount = pc->pointCount;
    Vec2 cA = (*positions)[indexA].c; float aA = (*positions)[indexA].a;
    Vec2 cB = (*positions)[indexB].c; float aB = (*positions)[indexB].a;
    for (int j = 0; j < pointCount; ++j) {
      Transform xfA, xfB;
      xfA.q.Set(aA);
      xfB.q.Set(aB);
      xfA.p = cA - Mul(xfA.q, localCenterA);
      xfB.p = cB - Mul(xfB.q, localCenterB);
      PositionSolverManifold psm;
      psm.Initialize(pc, xfA, xfB, j);
      Vec2 normal = psm.normal, point = psm.point;
      float separation = psm.separation;
      Vec2 rA = point - cA, rB = point - cB;
      minSeparation = Min(minSeparation, separation);
      float C = Clamp(kBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
      float rnA = Cross(rA, normal), rnB = Cross(rB, normal);
      float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      float impulse = K > 0.0f ? -C / K : 0.0f;
      Vec2 P = impulse * normal;
      cA -= mA * P;
      aA -= iA * Cross(rA, P);
      cB += mB * P;
      aB += iB * Cross(rB, P);
    }
    (*positions)[indexA].c = cA; (*positions)[indexA].a = aA;
    (*positions)[indexB].c = cB; (*positions)[indexB].a = aB;
  }
  return minSeparation >= -3.0f * kLinearSlop;
}

inline bool ContactSolver::SolveTOIPositionConstraints(int toiIndexA, int toiIndexB) {
  float minSeparation = 0.0f;
  for (size_t i = 0; i < pcs.size(); ++i) {
    ContactPositionConstraint* pc = &pcs[i];
    int indexA = pc->indexA, indexB = pc->indexB;
    Vec2 localCenterA = pc->localCenterA, localCenterB = pc->localCenterB;
    int pointCount = pc->pointCount;
    float mA = 0.0f, iA = 0.0f;
    if (indexA == toiIndexA || indexA == toiIndexB) { mA = pc->invMassA; iA = pc->invIA; }
    float mB = 0.0f, iB = 0.0f;
    if (indexB == toiIndexA || indexB == toiIndexB) { mB = pc->invMassB; iB = pc->invIB; }
    Vec2 cA = (*positions)[indexA].c; float aA = (*positions)[indexA].a;
    Vec2 cB = (*positions)[indexB].c; float aB = (*positions)[indexB].a;
    for (int j = 0; j < pointCount; ++j) {
      Transform xfA, xfB;
      xfA.q.Set(aA);
      xfB.q.Set(aB);
      xfA.p = cA - Mul(xfA.q, localCenterA);
      xfB.p = cB - Mul(xfB.q, localCenterB);
      PositionSolverManifold psm;
      psm.Initialize(pc, xfA, xfB, j);
      Vec2 normal = psm.normal, point = psm.point;
      float separation = psm.separation;
      Vec2 rA = point - cA, rB = point - cB;
      minSeparation = Min(minSeparation, separation);
      float C = Clamp(kToiBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
      float rnA = Cross(rA, normal), rnB = Cross(rB, normal);
      float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
      float impulse = K > 0.0f ? -C / K : 0.0f;
      Vec2 P = impulse * normal;
      cA -= mA * P;
      aA -= iA * Cross(rA, P);
      cB += mB * P;
      aB += iB * Cross(rB, P);
    }
    (*positions)[indexA].c = cA; (*positions)[indexA].a = aA;
    (*positions)[indexB].c = cB; (*positions)[indexB].a = aB;
  }
  return minSeparation >= -1.5f * kLinearSlop;
}

}  // namespace b2o
