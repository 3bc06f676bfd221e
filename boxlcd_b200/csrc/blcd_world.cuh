// blcd_world.cuh -- one simulated world per thread: WorldEnv.step's `b2World.Step` x3 (boxLCD/world_env.py:431-458),
// reset (world_env.py:197-385), _get_obs (world_env.py:387-429) and lcd_render (world_env.py:460-512) for a batch of
// independent worlds.
//
// Mapping: ONE THREAD PER WORLD.  Sequential impulses is a Gauss-Seidel sweep: every constraint reads the velocities the
// previous one wrote, and in these scenes all joints share the robot root, so there is no constraint-level parallelism
// to hand to a warp without changing the results.  Throughput comes from world-level parallelism; what the design has
// to get right is where the per-world working set lives:
//   * shared memory, laid out [word][thread] (stride = block size, so a warp touches 32 consecutive banks whatever
//     row each lane indexes): body velocities / positions / inverse masses and the velocity-constraint records of
//     joints and contacts -- everything the 180 velocity + <=60 position iterations per sub-step touch;
//   * HBM, laid out [word][world] (a warp reads one 128-byte line per word): the persistent state -- poses, velocities,
//     sleep timers, fat AABBs, joint warm-start impulses, the ordered contact list and the manifold slots;
//   * registers / local memory: everything transient (narrow phase, GJK, TOI bookkeeping, island DFS).
// The same code compiles for the host (tests/hostsim) so that its logic can be diffed bit-for-bit against the CPU oracle.
#pragma once
#include "blcd_collide.cuh"
#include "blcd_raster.cuh"

namespace BLCD_NS {

constexpr int kSceneBytes = (int)((sizeof(DScene) + 15) / 16 * 16);

#ifdef __CUDACC__
// dynamic shared memory of every simulation kernel: [DScene copy][hot words x block threads].  Going through this
// symbol (instead of a pointer stored in the Sim object) lets the compiler emit LDS/STS and know that these stores
// cannot alias the thread's local state.
extern __shared__ __align__(16) unsigned char blcd_smem[];
#endif

template <int STRIDE>
struct Hot {
  float* p;
#ifdef __CUDA_ARCH__
  BLCD_HD float* base() const { return reinterpret_cast<float*>(blcd_smem + kSceneBytes) + threadIdx.x; }
#else
  BLCD_HD float* base() const { return p; }
#endif
  BLCD_HD float& operator[](int i) const { return base()[i * STRIDE]; }
  BLCD_HD uint32_t& u(int i) const { return reinterpret_cast<uint32_t*>(base())[i * STRIDE]; }
};

struct Gw {  // view of one world's persistent words in HBM
  uint32_t* p;
  int64_t n;
  BLCD_HD uint32_t& u(int i) const { return p[(int64_t)i * n]; }
  BLCD_HD float& f(int i) const { return reinterpret_cast<float*>(p)[(int64_t)i * n]; }
};

// Philox4x32-10, counter = (block, world lo, world hi, 0), key = seed
struct Philox {
  uint32_t key0, key1, w0, w1, draws;
  uint32_t buf[4];
  uint32_t buf_block;
  BLCD_HD void init(uint64_t seed, uint64_t world, uint32_t d) {
    key0 = (uint32_t)seed; key1 = (uint32_t)(seed >> 32);
    w0 = (uint32_t)world; w1 = (uint32_t)(world >> 32);
    draws = d; buf_block = 0xFFFFFFFFu;
    buf[0] = buf[1] = buf[2] = buf[3] = 0u;
  }
  BLCD_HD void block(uint32_t blk) {
    uint32_t c0 = blk, c1 = w0, c2 = w1, c3 = 0u, k0 = key0, k1 = key1;
    for (int r = 0; r < 10; ++r) {
      uint64_t p0 = (uint64_t)0xD2511F53u * c0;
      uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
      uint32_t n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
      uint32_t n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    buf[0] = c0; buf[1] = c1; buf[2] = c2; buf[3] = c3;
    buf_block = blk;
  }
  BLCD_HD uint32_t next() {
    uint32_t blk = draws >> 2;
    if (blk != buf_block) block(blk);
    uint32_t lane = draws & 3u;
    ++draws;
    return lane == 0 ? buf[0] : (lane == 1 ? buf[1] : (lane == 2 ? buf[2] : buf[3]));
  }
  BLCD_HD double uniform(double lo, double hi) { return lo + (hi - lo) * ((double)next() * (1.0 / 4294967296.0)); }
};

// joint record in shared memory
enum { J_RAX = 0, J_RAY, J_RBX, J_RBY, J_EXX, J_EYX, J_EZX, J_EYY, J_EZY, J_EZZ, J_MM, J_IX, J_IY, J_IZ, J_MI, J_MS, J_PK, J_REF };
// contact record in shared memory
// C_PK packs: body row A (bits 0..4), row B (5..9), point count (10..11), manifold slot (12..19), island (20..)
enum { C_NX = 0, C_NY, C_FR, C_PK, C_K11, C_K12, C_K22, C_N11, C_N12, C_N22, C_PT };
enum { P_RAX = 0, P_RAY, P_RBX, P_RBY, P_NM, P_TM, P_BIAS, P_NI, P_TI };
// manifold slot in HBM
enum { S_HDR = 0, S_LNX, S_LNY, S_LPX, S_LPY, S_PT };  // point j at S_PT + 5 j: x, y, ni, ti, id
constexpr uint32_t kSlotFree = 0xFFFFFFFFu;

// Velocity-constraint record of one joint held in REGISTERS for the whole velocity loop (solve order k is a
// compile-time index after unrolling; only the body rows it touches are runtime values).  Everything that
// b2RevoluteJoint::SolveVelocityConstraints recomputes identically every iteration (cross(ey, ez), the 3x3 and 2x2
// determinants) is hoisted: same operations on the same inputs, so the results are bit-identical.
struct JV {
  int rowA, rowB, limitState, flags;  // flags: 1 = motor enabled, 2 = limit enabled, 4 = fixed rotation
  float rAx, rAy, rBx, rBy, exx, eyx, ezx, eyy, ezy, ezz;
  float cx, cy, cz, det3, det2, mm, maxImp, ms, mA, iA, mB, iB;
  float ix, iy, iz, mi;
};

template <int STRIDE>
struct Sim {
  const DScene* scene_host;
  Hot<STRIDE> hot;
#ifdef __CUDA_ARCH__
  BLCD_HD const DScene& scene() const { return *reinterpret_cast<const DScene*>(blcd_smem); }
#else
  BLCD_HD const DScene& scene() const { return *scene_host; }
#endif
#define sc scene()
  Gw g;
  // body state (registers / local memory)
  V2 c[kMaxBodies], c0[kMaxBodies], v[kMaxBodies];
  float a[kMaxBodies], a0[kMaxBodies], w[kMaxBodies], sleepT[kMaxBodies], alpha0[kMaxBodies];
  Xf xf[kMaxBodies];
  Box fat[kMaxBodies];
  float walpha0[BLCD_MAX_WALLS];
  uint32_t awake, variant, moved;
  bool newFixture;
  float inv_dt0;
  int32_t ep_t;
  Philox rng;
  // contacts
  uint8_t clist[kMaxPairs];
  int ncl;
  int8_t pslot[kMaxPairs];
  uint32_t slotUsed;
  uint32_t cnt[BLCD_N_COUNTERS];
  // velocity-constraint records.  They live in thread-local memory (L1-cached, [word][lane] interleaved by the hardware)
  // rather than in shared memory: only the records a warp actually uses occupy cache lines, so sixteen manifold slots
  // cost nothing until a world really has that many touching contacts, and shared memory is left for the body rows.
  float cr[kMaxSlots * kHotCon];
  float jr[kMaxJoints * kHotJoint];
  BLCD_HD uint32_t& cru(int i) { return reinterpret_cast<uint32_t*>(cr)[i]; }
  BLCD_HD uint32_t cru(int i) const { return reinterpret_cast<const uint32_t*>(cr)[i]; }
  BLCD_HD uint32_t& jru(int i) { return reinterpret_cast<uint32_t*>(jr)[i]; }
  BLCD_HD uint32_t jru(int i) const { return reinterpret_cast<const uint32_t*>(jr)[i]; }
  // islands and solve order (rebuilt every sub-step by solve_setup)
  int8_t islandOf[kMaxBodies];
  int nIslands, nc, njo;
  uint8_t jorder[kMaxJoints];   // joints in b2Island order
  uint8_t jisl[kMaxJoints];     // island of the k-th joint in solve order
  uint32_t islDone;             // islands whose position iterations converged (solve_position)

  // shared-memory row offsets and the static row index, copied out of the scene table once: the table itself sits in
  // shared memory, where every store to a hot row would force the compiler to re-read it
  int oV, oP, oM, nbS;
  uint32_t live;  // lanes of this warp that own a world (for warp re-convergence points)

  // world_end: one past the last world this LAUNCH covers (launches over a sub-range of the handle's worlds pass it)
  BLCD_HD Sim(const DScene& s, float* hot_base, uint32_t* state, int64_t n_worlds, int64_t world, int64_t world_end = -1) : scene_host(&s) {
    hot.p = hot_base;
    oV = sc.h_vel; oP = sc.h_pos; oM = sc.h_mass; nbS = sc.nb;
#ifdef __CUDA_ARCH__
    live = __activemask();
    {
      int64_t first = world - threadIdx.x;  // first world of this block
      int64_t left = (world_end < 0 ? n_worlds : world_end) - first;
      int warps = (int)((left + 31) / 32);
      int maxw = (int)(blockDim.x / 32);
      bar_threads = 32 * (warps < maxw ? warps : maxw);
    }
#else
    live = 1u;
    bar_threads = 0;
#endif
    g.p = state + world;
    g.n = n_worlds;
    x.p = nullptr;
    x.n = n_worlds;
  }
  BLCD_HD void attach_scratch(uint32_t* scratch, int64_t world) { x.p = scratch + world; }

#if defined(BLCD_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
  // diagnostic build: cycles per phase, accumulated into the counter words (which then no longer hold the usual counts)
  long long ph_last;
  BLCD_HD void ph_start() { ph_last = clock64(); }
  BLCD_HD void ph(int i) { long long t = clock64(); cnt[i] += (uint32_t)((t - ph_last) >> 6); ph_last = t; }
#else
  BLCD_HD void ph_start() {}
  BLCD_HD void ph(int) {}
#endif

  // lanes that took different numbers of TOI events / position iterations wait for each other here, so that the next
  // phase runs with the whole warp instead of as two half-empty groups chasing each other through the code
  BLCD_HD void reconverge() const {
#ifdef __CUDA_ARCH__
    __syncwarp(live);
#endif
  }

  // Phase alignment: all warps of the block enter the long loops (velocity / position iterations) together, so that
  // the SM's small instruction cache holds ONE loop body at a time instead of a different phase per warp.  ncu showed
  // the unaligned kernel limited by instruction fetch (sm__icc hit rate 69 %, gcc instruction requests 63 % of peak).
  // bar_threads = 32 x (warps of this block that own at least one world); every such warp reaches every phase point the
  // same number of times (T env steps x n_substeps), so the named barrier cannot deadlock.
  int bar_threads;
  BLCD_HD void phase_align(int finished_phase) {
#ifdef __CUDA_ARCH__
    __syncwarp(live);
#if defined(BLCD_PHASE_CLOCKS) && BLCD_PHASE_CLOCKS == 2
    ph(7);                // (diagnostic build 2) work is dumped, the WAIT is attributed to the phase that just ended
    asm volatile("bar.sync 1, %0;" ::"r"(bar_threads) : "memory");
    ph(finished_phase);
#else
    ph(finished_phase);   // (diagnostic build) work of the phase that just ended
    asm volatile("bar.sync 1, %0;" ::"r"(bar_threads) : "memory");
    ph(5);                // (diagnostic build) time spent waiting at the barrier
#endif
#endif
  }

  // ---- phase pipeline: spill / fill of what lives between the phases of one sub-step (blcd_pipeline.cuh) -------------
  Gw x;   // this world's scratch words in HBM
  BLCD_HD void x_rows_out(int first, int words) const {   // body rows: [v 3][c a 3][invMass invI 2] x nb, dynamic rows only
    const int nb = sc.nb;
    for (int b = 0; b < nb; ++b)
      for (int k = first; k < first + words; ++k)
        x.f(sc.x_rows + 8 * b + k) = k < 3 ? hot[oV + 3 * b + k] : (k < 6 ? hot[oP + 3 * b + k - 3] : hot[oM + 2 * b + k - 6]);
  }
  BLCD_HD void x_rows_in(int first, int words) const {
    const int nb = sc.nb;
    for (int b = 0; b < nb; ++b)
      for (int k = first; k < first + words; ++k) {
        float val = x.f(sc.x_rows + 8 * b + k);
        if (k < 3) hot[oV + 3 * b + k] = val; else if (k < 6) hot[oP + 3 * b + k - 3] = val; else hot[oM + 2 * b + k - 6] = val;
      }
    // the static row that stands in for every wall
    set_hv(nb, mk(0.0f, 0.0f), 0.0f);
    set_hc(nb, mk(0.0f, 0.0f), 0.0f);
    hot[oM + 2 * nb] = 0.0f;
    hot[oM + 2 * nb + 1] = 0.0f;
  }
  // lean variant for the velocity kernel, whose shared-memory column holds [v 3][invMass invI 2] per body and NO position
  // rows (the caller has set oM = oP): 5 instead of 8 words per body and thread, so more blocks fit an SM
  BLCD_HD void x_rows_in_lean() const {
    const int nb = sc.nb;
    for (int b = 0; b < nb; ++b) {
      for (int k = 0; k < 3; ++k) hot[oV + 3 * b + k] = x.f(sc.x_rows + 8 * b + k);
      for (int k = 0; k < 2; ++k) hot[oM + 2 * b + k] = x.f(sc.x_rows + 8 * b + 6 + k);
    }
    set_hv(nb, mk(0.0f, 0.0f), 0.0f);
    hot[oM + 2 * nb] = 0.0f;
    hot[oM + 2 * nb + 1] = 0.0f;
  }
  BLCD_HD void x_misc_out() const {
    x.u(sc.x_misc) = (uint32_t)nc | ((uint32_t)njo << 8) | ((uint32_t)nIslands << 16);
    x.u(sc.x_misc + 1) = islDone;
    const int jw = (sc.nj + 3) / 4, bw = (sc.nb + 3) / 4;
    for (int i = 0; i < jw; ++i) {
      uint32_t a_ = 0u, b_ = 0u;
      for (int k = 0; k < 4; ++k) {
        int j = 4 * i + k;
        if (j < njo) { a_ |= (uint32_t)jorder[j] << (8 * k); b_ |= (uint32_t)jisl[j] << (8 * k); }
      }
      x.u(sc.x_misc + 2 + i) = a_;
      x.u(sc.x_misc + 2 + jw + i) = b_;
    }
    for (int i = 0; i < bw; ++i) {
      uint32_t a_ = 0u;
      for (int k = 0; k < 4; ++k) {
        int b = 4 * i + k;
        if (b < sc.nb) a_ |= (uint32_t)(uint8_t)islandOf[b] << (8 * k);
      }
      x.u(sc.x_misc + 2 + 2 * jw + i) = a_;
    }
  }
  BLCD_HD void x_misc_in() {
    uint32_t w0 = x.u(sc.x_misc);
    nc = (int)(w0 & 255u); njo = (int)((w0 >> 8) & 255u); nIslands = (int)(w0 >> 16);
    islDone = x.u(sc.x_misc + 1);
    const int jw = (sc.nj + 3) / 4, bw = (sc.nb + 3) / 4;
    for (int i = 0; i < jw; ++i) {
      uint32_t a_ = x.u(sc.x_misc + 2 + i), b_ = x.u(sc.x_misc + 2 + jw + i);
      for (int k = 0; k < 4; ++k) {
        int j = 4 * i + k;
        if (j < kMaxJoints) { jorder[j] = (uint8_t)(a_ >> (8 * k)); jisl[j] = (uint8_t)(b_ >> (8 * k)); }
      }
    }
    for (int i = 0; i < bw; ++i) {
      uint32_t a_ = x.u(sc.x_misc + 2 + 2 * jw + i);
      for (int k = 0; k < 4; ++k) {
        int b = 4 * i + k;
        if (b < kMaxBodies) islandOf[b] = (int8_t)(uint8_t)(a_ >> (8 * k));
      }
    }
  }
  BLCD_HD void x_jr_out(int first, int words) const {
    for (int j = 0; j < sc.nj; ++j)
      for (int k = first; k < first + words; ++k) x.u(sc.x_jr + kHotJoint * j + k) = jru(kHotJoint * j + k);
  }
  BLCD_HD void x_jr_in(int first, int words) {
    for (int j = 0; j < sc.nj; ++j)
      for (int k = first; k < first + words; ++k) jru(kHotJoint * j + k) = x.u(sc.x_jr + kHotJoint * j + k);
  }
  // contact records: the header and as many point records as the solver uses (1 when the block solver was refused)
  BLCD_HD void x_cr_out() const {
    for (int k = 0; k < nc; ++k) {
      const int pts = (int)((cru(kHotCon * k + C_PK) >> 10) & 3u);
      const int words = kHotConHdr + kHotConPt * (pts < 1 ? 1 : pts);
      for (int i = 0; i < words; ++i) x.u(sc.x_cr + kHotCon * k + i) = cru(kHotCon * k + i);
    }
  }
  BLCD_HD void x_cr_in() {
    for (int k = 0; k < nc; ++k) {
      const uint32_t pk = x.u(sc.x_cr + kHotCon * k + C_PK);
      const int pts = (int)((pk >> 10) & 3u);
      const int words = kHotConHdr + kHotConPt * (pts < 1 ? 1 : pts);
      for (int i = 0; i < words; ++i) cru(kHotCon * k + i) = x.u(sc.x_cr + kHotCon * k + i);
    }
  }
  BLCD_HD void x_cr_pk_in() {   // position iterations only need the packed word (rows, slot, island)
    for (int k = 0; k < nc; ++k) cru(kHotCon * k + C_PK) = x.u(sc.x_cr + kHotCon * k + C_PK);
  }

  // ---- scene helpers ------------------------------------------------------------------------------------------------
  BLCD_HD int var_of(int b) const { return (int)((variant >> b) & 1u); }
  BLCD_HD const DShape& bshape(int b) const { return sc.body[b].shape[var_of(b)]; }
  BLCD_HD const DShape& fshape(int f) const { return f < sc.nw ? sc.wall[f] : bshape(f - sc.nw); }
  BLCD_HD V2 lc_of(int b) const { return sc.body[b].lc[var_of(b)]; }
  BLCD_HD int row_of(int f) const { return f < sc.nw ? sc.nb : f - sc.nw; }  // row nb = the static body
  BLCD_HD bool is_awake(int b) const { return (awake >> b) & 1u; }
  BLCD_HD void set_awake(int b) {  // b2Body::SetAwake(true)
    if (!is_awake(b)) { awake |= 1u << b; sleepT[b] = 0.0f; }
  }
  BLCD_HD void wake_fixture(int f) { if (f >= sc.nw) set_awake(f - sc.nw); }
  BLCD_HD Xf fxf(int f) const { return f < sc.nw ? xf_identity() : xf[f - sc.nw]; }
  BLCD_HD Box ffat(int f) const { return f < sc.nw ? sc.wallFat[f] : fat[f - sc.nw]; }
  // fixture order inside the contact after b2Contact::Create's type-ordering swap
  BLCD_HD void pair_ab(int p, int* fA, int* fB) const {
    int fa = sc.pair[p].fa, fb = sc.pair[p].fb;
    int ta = fshape(fa).type, tb = fshape(fb).type;
    bool primary = (ta == tb) || (ta == SH_POLY && tb == SH_CIRCLE) || (ta == SH_EDGE);
    *fA = primary ? fa : fb;
    *fB = primary ? fb : fa;
  }

  // ---- HBM <-> thread -----------------------------------------------------------------------------------------------
  BLCD_HD void load(uint64_t seed, int64_t global_world) {
    const int nb = sc.nb;
    uint32_t flags = g.u(sc.off_misc + 0);
    awake = flags & kAwakeMask;
    variant = kVariantInFlags ? ((flags >> kVariantShift) & kBodyMask) : g.u(sc.off_misc + 4);
    newFixture = (flags & kNewFixtureBit) != 0;
    inv_dt0 = g.f(sc.off_misc + 1);
    ep_t = (int32_t)g.u(sc.off_misc + 2);
    rng.init(seed, (uint64_t)global_world, g.u(sc.off_misc + 3));
    for (int b = 0; b < kMaxBodies; ++b) {
      if (b < nb) {
        int o = kBodyWords * b;
        c[b] = mk(g.f(o + 0), g.f(o + 1)); a[b] = g.f(o + 2);
        v[b] = mk(g.f(o + 3), g.f(o + 4)); w[b] = g.f(o + 5);
        sleepT[b] = g.f(o + 6);
        fat[b].lo = mk(g.f(o + 7), g.f(o + 8)); fat[b].hi = mk(g.f(o + 9), g.f(o + 10));
        c0[b] = c[b]; a0[b] = a[b]; alpha0[b] = 0.0f;
        // m_xf.p is kept alongside the sweep: after CreateBody / SetTransform it is the given position, which
        // c - q * localCenter reproduces only to an ulp when the local centre is not the origin (the luxo head)
        xf[b].q = rot_of(a[b]);
        xf[b].p = mk(g.f(o + 11), g.f(o + 12));
      }
    }
    for (int j = 0; j < sc.nj; ++j) {  // joint warm-start state goes straight into its shared-memory record
      int o = sc.off_joint + kJointWords * j, h = kHotJoint * j;
      jr[h + J_IX] = g.f(o + 0); jr[h + J_IY] = g.f(o + 1); jr[h + J_IZ] = g.f(o + 2);
      jr[h + J_MI] = g.f(o + 3); jr[h + J_MS] = g.f(o + 4);
      jru(h + J_PK) = g.u(o + 5) & 3u;  // limit state; solve-order fields are filled per sub-step
      jr[h + J_REF] = g.f(o + 6);
    }
    ncl = (int)g.u(sc.off_clist);
    for (int k = 0; k < kMaxPairs; ++k) pslot[k] = -1;
    for (int k = 0; k < ncl; ++k) clist[k] = (uint8_t)((g.u(sc.off_clist + 1 + (k >> 2)) >> (8 * (k & 3))) & 0xFFu);
    slotUsed = 0u;
    for (int s = 0; s < sc.maxm; ++s) {
      uint32_t hdr = g.u(sc.off_slots + kSlotWords * s + S_HDR);
      if (hdr != kSlotFree) { slotUsed |= 1u << s; pslot[hdr & 0xFFu] = (int8_t)s; }
    }
    for (int k = 0; k < BLCD_N_COUNTERS; ++k) cnt[k] = g.u(sc.off_cnt + k);
    moved = 0u;
  }

  BLCD_HD void load_variant() {   // shape variants only (pipeline phases that need local centres / radii but no body state)
    variant = kVariantInFlags ? ((g.u(sc.off_misc + 0) >> kVariantShift) & kBodyMask) : g.u(sc.off_misc + 4);
  }

  // what the setup phase of the pipeline can have changed: flags (awake, new-fixture), sleep timers, the contact list, the
  // motor speeds, counters, episode step and the draw counter -- not the body poses / velocities / AABBs, not the joint
  // impulses (manifold slots are always written in place)
  BLCD_HD void store_after_setup() {
    const int nb = sc.nb;
    if (kVariantInFlags) g.u(sc.off_misc + 0) = (awake & kAwakeMask) | (variant << kVariantShift) | (newFixture ? kNewFixtureBit : 0u);
    else g.u(sc.off_misc + 0) = (awake & kAwakeMask) | (newFixture ? kNewFixtureBit : 0u);
    g.u(sc.off_misc + 2) = (uint32_t)ep_t;
    g.u(sc.off_misc + 3) = rng.draws;
    for (int b = 0; b < kMaxBodies; ++b)
      if (b < nb) g.f(kBodyWords * b + 6) = sleepT[b];
    for (int j = 0; j < sc.nj; ++j) g.f(sc.off_joint + kJointWords * j + 4) = jr[kHotJoint * j + J_MS];
    store_clist();
    for (int k = 0; k < BLCD_N_COUNTERS; ++k) g.u(sc.off_cnt + k) = cnt[k];
  }

  BLCD_HD void store_clist() {
    g.u(sc.off_clist) = (uint32_t)ncl;
    for (int k4 = 0; k4 < sc.clist_words - 1; ++k4) {
      uint32_t wd = 0u;
      for (int k = 0; k < 4; ++k) {
        int i = 4 * k4 + k;
        if (i < ncl) wd |= (uint32_t)clist[i] << (8 * k);
      }
      g.u(sc.off_clist + 1 + k4) = wd;
    }
  }

  BLCD_HD void store() {
    const int nb = sc.nb;
    if (kVariantInFlags) {
      g.u(sc.off_misc + 0) = (awake & kAwakeMask) | (variant << kVariantShift) | (newFixture ? kNewFixtureBit : 0u);
    } else {
      g.u(sc.off_misc + 0) = (awake & kAwakeMask) | (newFixture ? kNewFixtureBit : 0u);
      g.u(sc.off_misc + 4) = variant;
    }
    g.f(sc.off_misc + 1) = inv_dt0;
    g.u(sc.off_misc + 2) = (uint32_t)ep_t;
    g.u(sc.off_misc + 3) = rng.draws;
    for (int b = 0; b < kMaxBodies; ++b) {
      if (b < nb) {
        int o = kBodyWords * b;
        g.f(o + 0) = c[b].x; g.f(o + 1) = c[b].y; g.f(o + 2) = a[b];
        g.f(o + 3) = v[b].x; g.f(o + 4) = v[b].y; g.f(o + 5) = w[b];
        g.f(o + 6) = sleepT[b];
        g.f(o + 7) = fat[b].lo.x; g.f(o + 8) = fat[b].lo.y; g.f(o + 9) = fat[b].hi.x; g.f(o + 10) = fat[b].hi.y;
        g.f(o + 11) = xf[b].p.x; g.f(o + 12) = xf[b].p.y;
      }
    }
    for (int j = 0; j < sc.nj; ++j) {
      int o = sc.off_joint + kJointWords * j, h = kHotJoint * j;
      g.f(o + 0) = jr[h + J_IX]; g.f(o + 1) = jr[h + J_IY]; g.f(o + 2) = jr[h + J_IZ];
      g.f(o + 3) = jr[h + J_MI]; g.f(o + 4) = jr[h + J_MS];
      g.u(o + 5) = jru(h + J_PK) & 3u;
      g.f(o + 6) = jr[h + J_REF];
    }
    g.u(sc.off_clist) = (uint32_t)ncl;
    for (int k4 = 0; k4 < sc.clist_words - 1; ++k4) {
      uint32_t wd = 0u;
      for (int k = 0; k < 4; ++k) {
        int i = 4 * k4 + k;
        if (i < ncl) wd |= (uint32_t)clist[i] << (8 * k);
      }
      g.u(sc.off_clist + 1 + k4) = wd;
    }
    for (int k = 0; k < BLCD_N_COUNTERS; ++k) g.u(sc.off_cnt + k) = cnt[k];
  }

  // ---- manifold slots -----------------------------------------------------------------------------------------------
  BLCD_HD int slot_base(int s) const { return sc.off_slots + kSlotWords * s; }
  BLCD_HD void slot_write(int s, int p, const Mf& m) {
    int o = slot_base(s);
    g.u(o + S_HDR) = (uint32_t)p | ((uint32_t)m.type << 8) | ((uint32_t)m.count << 16);
    g.f(o + S_LNX) = m.ln.x; g.f(o + S_LNY) = m.ln.y; g.f(o + S_LPX) = m.lp.x; g.f(o + S_LPY) = m.lp.y;
    for (int j = 0; j < 2; ++j) {
      if (j < m.count) {
        int q = o + S_PT + 5 * j;
        g.f(q + 0) = m.pt[j].x; g.f(q + 1) = m.pt[j].y; g.f(q + 2) = m.ni[j]; g.f(q + 3) = m.ti[j]; g.u(q + 4) = m.id[j];
      }
    }
  }

  // b2Contact::Update for pair p: re-evaluate the manifold at the current transforms, carry warm-start impulses over
  // by contact id, keep the slot table in step with `touching`.  Returns touching.
  BLCD_HDN bool update_contact(int p) {
    int fA, fB;
    pair_ab(p, &fA, &fB);
    const DShape &A = fshape(fA), &B = fshape(fB);
    Mf m;
    m.count = 0; m.type = 0; m.ln = mk(0.0f, 0.0f); m.lp = mk(0.0f, 0.0f);
    if (A.type == SH_EDGE) {
      if (B.type == SH_CIRCLE) collide_edge_circle(m, A, B, fxf(fB));
      else collide_edge_polygon(m, A, B, fxf(fB));
    } else if (A.type == SH_CIRCLE) {
      collide_circles(m, A, fxf(fA), B, fxf(fB));
    } else if (B.type == SH_CIRCLE) {
      collide_polygon_circle(m, A, fxf(fA), B, fxf(fB));
    } else {
      collide_polygons(m, A, fxf(fA), B, fxf(fB), (sc.flags & BLCD_FLAG_REFFACE_2_3_0) != 0);
    }
    int s = pslot[p];
    bool wasTouching = s >= 0;
    for (int i = 0; i < 2; ++i) {
      if (i < m.count) {
        m.ni[i] = 0.0f; m.ti[i] = 0.0f;
        if (wasTouching) {
          int o = slot_base(s);
          int oldCount = (int)((g.u(o + S_HDR) >> 16) & 0xFFu);
          for (int j = 0; j < oldCount; ++j) {
            int q = o + S_PT + 5 * j;
            if (g.u(q + 4) == m.id[i]) { m.ni[i] = g.f(q + 2); m.ti[i] = g.f(q + 3); break; }
          }
        }
      }
    }
    bool touching = m.count > 0;
    if (touching && s < 0) {  // allocate
      uint32_t freeMask = ~slotUsed & ((sc.maxm >= 32) ? 0xFFFFFFFFu : ((1u << sc.maxm) - 1u));
      if (freeMask == 0u) { ++cnt[BLCD_CNT_OVERFLOW]; touching = false; }
      else {
        s = 0;
        while (!((freeMask >> s) & 1u)) ++s;
        slotUsed |= 1u << s;
        pslot[p] = (int8_t)s;
      }
    }
    if (touching) slot_write(s, p, m);
    else if (s >= 0) {  // release
      g.u(slot_base(s) + S_HDR) = kSlotFree;
      slotUsed &= ~(1u << s);
      pslot[p] = -1;
    }
    if (touching != wasTouching) { wake_fixture(fA); wake_fixture(fB); }
    return touching;
  }

  // ---- broad phase bookkeeping ---------------------------------------------------------------------------------------
  // b2BroadPhase::UpdatePairs + b2ContactManager::AddPair: candidate pairs are stored in ascending proxy order and each
  // new contact is pushed to the head of the list, as Box2D does after sorting its pair buffer.
  BLCD_HDN void find_new_contacts(uint32_t movedMask) {
    if (movedMask == 0u) return;
    PairMask exists = PairMask::none();
    for (int k = 0; k < ncl; ++k) exists.set(clist[k]);
    for (int p = 0; p < sc.np; ++p) {
      int fa = sc.pair[p].fa, fb = sc.pair[p].fb;
      if (!(((movedMask >> fa) | (movedMask >> fb)) & 1u)) continue;
      if (exists.test(p)) continue;
      if (!box_overlap(ffat(fa), ffat(fb))) continue;
      for (int k = ncl; k > 0; --k) clist[k] = clist[k - 1];
      clist[0] = (uint8_t)p;
      ++ncl;
      wake_fixture(fa);
      wake_fixture(fb);
    }
  }

  // b2Fixture::Synchronize + b2DynamicTree::MoveProxy for body b moving from xf1 to its current transform
  BLCD_HD void sync_fixture(int b, const Xf& xf1) {
    const DShape& s = bshape(b);
    Box b1 = shape_aabb(s, xf1), b2 = shape_aabb(s, xf[b]), bb;
    bb.lo = mk(fminb(b1.lo.x, b2.lo.x), fminb(b1.lo.y, b2.lo.y));
    bb.hi = mk(fmaxb(b1.hi.x, b2.hi.x), fmaxb(b1.hi.y, b2.hi.y));
    if (box_contains(fat[b], bb)) return;
    V2 disp = xf[b].p - xf1.p;
    V2 r = mk(kAabbExtension, kAabbExtension);
    Box f;
    f.lo = bb.lo - r;
    f.hi = bb.hi + r;
    V2 d = kAabbMultiplier * disp;
    if (d.x < 0.0f) f.lo.x += d.x; else f.hi.x += d.x;
    if (d.y < 0.0f) f.lo.y += d.y; else f.hi.y += d.y;
    fat[b] = f;
    moved |= 1u << (sc.nw + b);
  }

  // b2ContactManager::Collide
  BLCD_HDN void collide() {
    int k = 0;
    while (k < ncl) {
      int p = clist[k];
      int fa = sc.pair[p].fa, fb = sc.pair[p].fb;
      bool activeA = fa >= sc.nw && is_awake(fa - sc.nw);
      bool activeB = fb >= sc.nw && is_awake(fb - sc.nw);
      if (!activeA && !activeB) { ++k; continue; }
      if (!box_overlap(ffat(fa), ffat(fb))) {  // b2ContactManager::Destroy
        int s = pslot[p];
        if (s >= 0) {
          wake_fixture(fa); wake_fixture(fb);
          g.u(slot_base(s) + S_HDR) = kSlotFree;
          slotUsed &= ~(1u << s);
          pslot[p] = -1;
        }
        for (int i = k; i + 1 < ncl; ++i) clist[i] = clist[i + 1];
        --ncl;
        continue;
      }
      update_contact(p);
      ++k;
    }
  }

  // ---- solver: shared-memory rows ------------------------------------------------------------------------------------
  BLCD_HD V2 hv(int r) const { return mk(hot[oV + 3 * r], hot[oV + 3 * r + 1]); }
  BLCD_HD float hw(int r) const { return hot[oV + 3 * r + 2]; }
  BLCD_HD void set_hv(int r, V2 x, float ww) const { hot[oV + 3 * r] = x.x; hot[oV + 3 * r + 1] = x.y; hot[oV + 3 * r + 2] = ww; }
  BLCD_HD V2 hc(int r) const { return mk(hot[oP + 3 * r], hot[oP + 3 * r + 1]); }
  BLCD_HD float ha(int r) const { return hot[oP + 3 * r + 2]; }
  BLCD_HD void set_hc(int r, V2 x, float aa) const { hot[oP + 3 * r] = x.x; hot[oP + 3 * r + 1] = x.y; hot[oP + 3 * r + 2] = aa; }
  BLCD_HD float hm(int r) const { return hot[oM + 2 * r]; }
  BLCD_HD float hi(int r) const { return hot[oM + 2 * r + 1]; }
  BLCD_HD V2 row_lc(int r) const { return r < nbS ? lc_of(r) : mk(0.0f, 0.0f); }

  BLCD_HD void stage_rows() {
    for (int b = 0; b < sc.nb; ++b) {
      set_hv(b, v[b], w[b]);
      set_hc(b, c[b], a[b]);
      hot[oM + 2 * b] = sc.body[b].invMass[var_of(b)];
      hot[oM + 2 * b + 1] = sc.body[b].invI[var_of(b)];
    }
    set_hv(sc.nb, mk(0.0f, 0.0f), 0.0f);
    set_hc(sc.nb, mk(0.0f, 0.0f), 0.0f);
    hot[oM + 2 * sc.nb] = 0.0f;
    hot[oM + 2 * sc.nb + 1] = 0.0f;
  }

  // b2ContactSolver ctor + InitializeVelocityConstraints for the manifold in slot s -> contact record k
  BLCD_HDN void contact_init(int k, int s, int isl, float dtRatio, bool warm) {
    const int o = slot_base(s), h = kHotCon * k;
    uint32_t hdr = g.u(o + S_HDR);
    int p = (int)(hdr & 0xFFu), type = (int)((hdr >> 8) & 0xFFu), count = (int)((hdr >> 16) & 0xFFu);
    int fA, fB;
    pair_ab(p, &fA, &fB);
    int rA_ = row_of(fA), rB_ = row_of(fB);
    float radiusA = fshape(fA).radius, radiusB = fshape(fB).radius;
    float frA = fA < sc.nw ? 0.2f : sc.body[fA - sc.nw].friction, frB = fB < sc.nw ? 0.2f : sc.body[fB - sc.nw].friction;
    float reA = fA < sc.nw ? 0.0f : sc.body[fA - sc.nw].restitution, reB = fB < sc.nw ? 0.0f : sc.body[fB - sc.nw].restitution;
    float friction = sqrtf(frA * frB), restitution = reA > reB ? reA : reB;
    float mA = hm(rA_), iA = hi(rA_), mB = hm(rB_), iB = hi(rB_);
    V2 cA = hc(rA_), cB = hc(rB_);
    float aA = ha(rA_), aB = ha(rB_);
    V2 vA = hv(rA_), vB = hv(rB_);
    float wA = hw(rA_), wB = hw(rB_);
    Xf xfA, xfB;
    xfA.q = rA_ < nbS ? rot_of(aA) : rot_identity();
    xfB.q = rB_ < nbS ? rot_of(aB) : rot_identity();
    xfA.p = cA - rmul(xfA.q, row_lc(rA_));
    xfB.p = cB - rmul(xfB.q, row_lc(rB_));
    V2 ln = mk(g.f(o + S_LNX), g.f(o + S_LNY)), lp = mk(g.f(o + S_LPX), g.f(o + S_LPY));
    V2 normal = mk(1.0f, 0.0f);
    V2 wp[2];
    wp[0] = wp[1] = mk(0.0f, 0.0f);
    // b2WorldManifold::Initialize
    if (type == MF_CIRCLES) {
      V2 pointA = xmul(xfA, lp);
      V2 pointB = xmul(xfB, mk(g.f(o + S_PT), g.f(o + S_PT + 1)));
      if (dist2(pointA, pointB) > kEps * kEps) { normal = pointB - pointA; normalize(normal); }
      V2 ca = pointA + radiusA * normal;
      V2 cb = pointB - radiusB * normal;
      wp[0] = 0.5f * (ca + cb);
    } else if (type == MF_FACE_A) {
      normal = rmul(xfA.q, ln);
      V2 planePoint = xmul(xfA, lp);
      for (int j = 0; j < 2; ++j) {
        if (j < count) {
          V2 clip = xmul(xfB, mk(g.f(o + S_PT + 5 * j), g.f(o + S_PT + 5 * j + 1)));
          V2 ca = clip + (radiusA - dot(clip - planePoint, normal)) * normal;
          V2 cb = clip - radiusB * normal;
          wp[j] = 0.5f * (ca + cb);
        }
      }
    } else {
      normal = rmul(xfB.q, ln);
      V2 planePoint = xmul(xfB, lp);
      for (int j = 0; j < 2; ++j) {
        if (j < count) {
          V2 clip = xmul(xfA, mk(g.f(o + S_PT + 5 * j), g.f(o + S_PT + 5 * j + 1)));
          V2 cb = clip + (radiusB - dot(clip - planePoint, normal)) * normal;
          V2 ca = clip - radiusA * normal;
          wp[j] = 0.5f * (ca + cb);
        }
      }
      normal = -normal;
    }
    cr[h + C_NX] = normal.x; cr[h + C_NY] = normal.y; cr[h + C_FR] = friction;
    V2 tangent = cross(normal, 1.0f);
    V2 prA[2], prB[2];
    for (int j = 0; j < 2; ++j) {
      if (j < count) {
        int q = h + C_PT + kHotConPt * j;
        V2 rA = wp[j] - cA, rB = wp[j] - cB;
        prA[j] = rA; prB[j] = rB;
        cr[q + P_RAX] = rA.x; cr[q + P_RAY] = rA.y; cr[q + P_RBX] = rB.x; cr[q + P_RBY] = rB.y;
        float rnA = cross(rA, normal), rnB = cross(rB, normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        cr[q + P_NM] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
        float rtA = cross(rA, tangent), rtB = cross(rB, tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        cr[q + P_TM] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
        float bias = 0.0f;
        float vRel = dot(normal, vB + cross(wB, rB) - vA - cross(wA, rA));
        if (vRel < -kVelocityThreshold) bias = -restitution * vRel;
        cr[q + P_BIAS] = bias;
        cr[q + P_NI] = warm ? dtRatio * g.f(o + S_PT + 5 * j + 2) : 0.0f;
        cr[q + P_TI] = warm ? dtRatio * g.f(o + S_PT + 5 * j + 3) : 0.0f;
      }
    }
    int pointCount = count;
    if (count == 2) {
      float rn1A = cross(prA[0], normal), rn1B = cross(prB[0], normal);
      float rn2A = cross(prA[1], normal), rn2B = cross(prB[1], normal);
      float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
      float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
      float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
      if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
        cr[h + C_K11] = k11; cr[h + C_K12] = k12; cr[h + C_K22] = k22;
        float det = k11 * k22 - k12 * k12;
        if (det != 0.0f) det = 1.0f / det;
        cr[h + C_N11] = det * k22; cr[h + C_N12] = -det * k12; cr[h + C_N22] = det * k11;
      } else {
        pointCount = 1;
      }
    }
    cru(h + C_PK) = (uint32_t)rA_ | ((uint32_t)rB_ << 5) | ((uint32_t)pointCount << 10) | ((uint32_t)s << 12) | ((uint32_t)isl << 20);
  }

  BLCD_HD void contact_warm_start(int k) {
    const int h = kHotCon * k;
    uint32_t pk = cru(h + C_PK);
    int rA_ = pk & 31u, rB_ = (pk >> 5) & 31u, count = (pk >> 10) & 3u;
    float mA = hm(rA_), iA = hi(rA_), mB = hm(rB_), iB = hi(rB_);
    V2 vA = hv(rA_), vB = hv(rB_);
    float wA = hw(rA_), wB = hw(rB_);
    V2 normal = mk(cr[h + C_NX], cr[h + C_NY]);
    V2 tangent = cross(normal, 1.0f);
    for (int j = 0; j < 2; ++j) {
      if (j < count) {
        int q = h + C_PT + kHotConPt * j;
        V2 rA = mk(cr[q + P_RAX], cr[q + P_RAY]), rB = mk(cr[q + P_RBX], cr[q + P_RBY]);
        V2 P = cr[q + P_NI] * normal + cr[q + P_TI] * tangent;
        wA -= iA * cross(rA, P);
        vA -= mA * P;
        wB += iB * cross(rB, P);
        vB += mB * P;
      }
    }
    set_hv(rA_, vA, wA);
    set_hv(rB_, vB, wB);
  }

  // returns true if any impulse applied by this sweep was non-zero
  BLCD_HD bool contact_solve_velocity(int k) { return contact_solve_velocity_rec(cr + kHotCon * k); }

  // the solve on a record given by address: r points into cr[] (above) or at a register copy of one record (the streamed
  // loop of solve_velocity); every index below is a compile-time constant once the two-point loops are unrolled
  BLCD_HD static uint32_t rec_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
  }
  // register copy of the k-th record (both points' words: which are live is only known from the record itself) and the
  // write-back of the four impulse words a solve may change
  BLCD_HD void rec_fetch(float (&r)[kHotCon], int k) const {
    const float* p = cr + kHotCon * k;
#pragma unroll
    for (int i = 0; i < kHotCon; ++i) r[i] = p[i];
  }
  BLCD_HD void rec_save(const float (&r)[kHotCon], int k) {
    float* p = cr + kHotCon * k;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      p[C_PT + kHotConPt * j + P_NI] = r[C_PT + kHotConPt * j + P_NI];
      p[C_PT + kHotConPt * j + P_TI] = r[C_PT + kHotConPt * j + P_TI];
    }
  }
  BLCD_HD bool contact_solve_velocity_rec(float* r) {
    uint32_t pk = rec_bits(r[C_PK]);
    int rA_ = pk & 31u, rB_ = (pk >> 5) & 31u, count = (pk >> 10) & 3u;
    if (rA_ == nbS) return contact_solve_velocity_wall(r, rB_, count);
    float mA = hm(rA_), iA = hi(rA_), mB = hm(rB_), iB = hi(rB_);
    V2 vA = hv(rA_), vB = hv(rB_);
    float wA = hw(rA_), wB = hw(rB_);
    V2 normal = mk(r[C_NX], r[C_NY]);
    V2 tangent = cross(normal, 1.0f);
    float friction = r[C_FR];
    bool changed = false;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j < count) {
        const int q = C_PT + kHotConPt * j;
        V2 rA = mk(r[q + P_RAX], r[q + P_RAY]), rB = mk(r[q + P_RBX], r[q + P_RBY]);
        V2 dv = vB + cross(wB, rB) - vA - cross(wA, rA);
        float vt = dot(dv, tangent) - 0.0f;
        float lambda = r[q + P_TM] * (-vt);
        float maxFriction = friction * r[q + P_NI];
        float oldImpulse = r[q + P_TI];
        float newImpulse = clampb(oldImpulse + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - oldImpulse;
        changed |= (lambda != 0.0f);
        r[q + P_TI] = newImpulse;
        V2 P = lambda * tangent;
        vA -= mA * P;
        wA -= iA * cross(rA, P);
        vB += mB * P;
        wB += iB * cross(rB, P);
      }
    }
    if (count == 1) {
      const int q = C_PT;
      V2 rA = mk(r[q + P_RAX], r[q + P_RAY]), rB = mk(r[q + P_RBX], r[q + P_RBY]);
      V2 dv = vB + cross(wB, rB) - vA - cross(wA, rA);
      float vn = dot(dv, normal);
      float lambda = -r[q + P_NM] * (vn - r[q + P_BIAS]);
      float oldImpulse = r[q + P_NI];
      float newImpulse = fmaxb(oldImpulse + lambda, 0.0f);
      lambda = newImpulse - oldImpulse;
      changed |= (lambda != 0.0f);
      r[q + P_NI] = newImpulse;
      V2 P = lambda * normal;
      vA -= mA * P;
      wA -= iA * cross(rA, P);
      vB += mB * P;
      wB += iB * cross(rB, P);
    } else if (count == 2) {
      const int q1 = C_PT, q2 = C_PT + kHotConPt;
      V2 r1A = mk(r[q1 + P_RAX], r[q1 + P_RAY]), r1B = mk(r[q1 + P_RBX], r[q1 + P_RBY]);
      V2 r2A = mk(r[q2 + P_RAX], r[q2 + P_RAY]), r2B = mk(r[q2 + P_RBX], r[q2 + P_RBY]);
      V2 aa = mk(r[q1 + P_NI], r[q2 + P_NI]);
      V2 dv1 = vB + cross(wB, r1B) - vA - cross(wA, r1A);
      V2 dv2 = vB + cross(wB, r2B) - vA - cross(wA, r2A);
      float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
      V2 b = mk(vn1 - r[q1 + P_BIAS], vn2 - r[q2 + P_BIAS]);
      V2 x;
      bool found = block_solve(r, q1, q2, aa, b, x);
      if (found) {
        V2 d = x - aa;
        changed |= (d.x != 0.0f) | (d.y != 0.0f);
        V2 P1 = d.x * normal, P2 = d.y * normal;
        vA -= mA * (P1 + P2);
        wA -= iA * (cross(r1A, P1) + cross(r2A, P2));
        vB += mB * (P1 + P2);
        wB += iB * (cross(r1B, P1) + cross(r2B, P2));
        r[q1 + P_NI] = x.x;
        r[q2 + P_NI] = x.y;
      }
    }
    set_hv(rA_, vA, wA);
    set_hv(rB_, vB, wB);
    return changed;
  }

  // b2ContactSolver's 2-point block LCP (cases 1-4); b is vn - bias, already reduced by K a inside
  BLCD_HD bool block_solve(const float* r, int q1, int q2, V2 aa, V2 b, V2& x) const {
    float k11 = r[C_K11], k12 = r[C_K12], k22 = r[C_K22];
    b -= mk(k11 * aa.x + k12 * aa.y, k12 * aa.x + k22 * aa.y);
    float n11 = r[C_N11], n12 = r[C_N12], n22 = r[C_N22];
    x = -mk(n11 * b.x + n12 * b.y, n12 * b.x + n22 * b.y);
    if (x.x >= 0.0f && x.y >= 0.0f) return true;
    x.x = -r[q1 + P_NM] * b.x; x.y = 0.0f;
    float vn2 = k12 * x.x + b.y;
    if (x.x >= 0.0f && vn2 >= 0.0f) return true;
    x.x = 0.0f; x.y = -r[q2 + P_NM] * b.y;
    float vn1 = k12 * x.y + b.x;
    if (x.y >= 0.0f && vn1 >= 0.0f) return true;
    x.x = 0.0f; x.y = 0.0f;
    return b.x >= 0.0f && b.y >= 0.0f;
  }

  // contact against a wall: body A is the static row, whose velocity and inverse masses are exactly zero, so every
  // A-side term of b2ContactSolver::SolveVelocityConstraints is an exact +-0 and is skipped
  BLCD_HD bool contact_solve_velocity_wall(float* r, int rB_, int count) {
    float mB = hm(rB_), iB = hi(rB_);
    V2 vB = hv(rB_);
    float wB = hw(rB_);
    V2 normal = mk(r[C_NX], r[C_NY]);
    V2 tangent = cross(normal, 1.0f);
    float friction = r[C_FR];
    bool changed = false;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j < count) {
        const int q = C_PT + kHotConPt * j;
        V2 rB = mk(r[q + P_RBX], r[q + P_RBY]);
        V2 dv = vB + cross(wB, rB);
        float vt = dot(dv, tangent) - 0.0f;
        float lambda = r[q + P_TM] * (-vt);
        float maxFriction = friction * r[q + P_NI];
        float oldImpulse = r[q + P_TI];
        float newImpulse = clampb(oldImpulse + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - oldImpulse;
        changed |= (lambda != 0.0f);
        r[q + P_TI] = newImpulse;
        V2 P = lambda * tangent;
        vB += mB * P;
        wB += iB * cross(rB, P);
      }
    }
    if (count == 1) {
      const int q = C_PT;
      V2 rB = mk(r[q + P_RBX], r[q + P_RBY]);
      V2 dv = vB + cross(wB, rB);
      float vn = dot(dv, normal);
      float lambda = -r[q + P_NM] * (vn - r[q + P_BIAS]);
      float oldImpulse = r[q + P_NI];
      float newImpulse = fmaxb(oldImpulse + lambda, 0.0f);
      lambda = newImpulse - oldImpulse;
      changed |= (lambda != 0.0f);
      r[q + P_NI] = newImpulse;
      V2 P = lambda * normal;
      vB += mB * P;
      wB += iB * cross(rB, P);
    } else if (count == 2) {
      const int q1 = C_PT, q2 = C_PT + kHotConPt;
      V2 r1B = mk(r[q1 + P_RBX], r[q1 + P_RBY]), r2B = mk(r[q2 + P_RBX], r[q2 + P_RBY]);
      V2 aa = mk(r[q1 + P_NI], r[q2 + P_NI]);
      V2 dv1 = vB + cross(wB, r1B);
      V2 dv2 = vB + cross(wB, r2B);
      float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
      V2 b = mk(vn1 - r[q1 + P_BIAS], vn2 - r[q2 + P_BIAS]);
      V2 x;
      bool found = block_solve(r, q1, q2, aa, b, x);
      if (found) {
        V2 d = x - aa;
        changed |= (d.x != 0.0f) | (d.y != 0.0f);
        V2 P1 = d.x * normal, P2 = d.y * normal;
        vB += mB * (P1 + P2);
        wB += iB * (cross(r1B, P1) + cross(r2B, P2));
        r[q1 + P_NI] = x.x;
        r[q2 + P_NI] = x.y;
      }
    }
    set_hv(rB_, vB, wB);
    return changed;
  }

  BLCD_HD void contact_store_impulses(int k) {
    const int h = kHotCon * k;
    uint32_t pk = cru(h + C_PK);
    int count = (pk >> 10) & 3u, s = (pk >> 12) & 255u;
    int o = slot_base(s);
    for (int j = 0; j < 2; ++j) {
      if (j < count) {
        int q = h + C_PT + kHotConPt * j;
        g.f(o + S_PT + 5 * j + 2) = cr[q + P_NI];
        g.f(o + S_PT + 5 * j + 3) = cr[q + P_TI];
      }
    }
  }

  // one pass of b2ContactSolver::SolvePositionConstraints / SolveTOIPositionConstraints over contact record k;
  // returns the smallest separation seen
  // position-solve copy of a manifold inside the contact record (the words the velocity solve uses for its point records):
  // the pipeline's position kernel reads the slot once per world instead of once per sweep (pos_cache_manifold)
  enum { M_HDR = C_PT, M_LNX, M_LNY, M_LPX, M_LPY, M_P0X, M_P0Y, M_P1X, M_P1Y };
  BLCD_HD void pos_cache_manifold(int k) {
    const int h = kHotCon * k;
    const int o = slot_base((int)((cru(h + C_PK) >> 12) & 255u));
    cru(h + M_HDR) = g.u(o + S_HDR);
    cr[h + M_LNX] = g.f(o + S_LNX); cr[h + M_LNY] = g.f(o + S_LNY); cr[h + M_LPX] = g.f(o + S_LPX); cr[h + M_LPY] = g.f(o + S_LPY);
    cr[h + M_P0X] = g.f(o + S_PT); cr[h + M_P0Y] = g.f(o + S_PT + 1); cr[h + M_P1X] = g.f(o + S_PT + 5); cr[h + M_P1Y] = g.f(o + S_PT + 6);
  }

  BLCD_HD float contact_solve_position(int k, float baumgarte) { return contact_solve_position_t<false>(k, baumgarte); }

  template <bool CACHED>
  BLCD_HD float contact_solve_position_t(int k, float baumgarte) {
    const int h = kHotCon * k;
    uint32_t pk = cru(h + C_PK);
    int rA_ = pk & 31u, rB_ = (pk >> 5) & 31u, s = (pk >> 12) & 255u;
    const int o = slot_base(s);
    uint32_t hdr = CACHED ? cru(h + M_HDR) : g.u(o + S_HDR);
    int p = (int)(hdr & 0xFFu), type = (int)((hdr >> 8) & 0xFFu), count = (int)((hdr >> 16) & 0xFFu);
    int fA, fB;
    pair_ab(p, &fA, &fB);
    float radiusA = fshape(fA).radius, radiusB = fshape(fB).radius;
    V2 lcA = row_lc(rA_), lcB = row_lc(rB_);
    float mA = hm(rA_), iA = hi(rA_), mB = hm(rB_), iB = hi(rB_);
    V2 cA = hc(rA_), cB = hc(rB_);
    float aA = ha(rA_), aB = ha(rB_);
    V2 ln = CACHED ? mk(cr[h + M_LNX], cr[h + M_LNY]) : mk(g.f(o + S_LNX), g.f(o + S_LNY));
    V2 lp = CACHED ? mk(cr[h + M_LPX], cr[h + M_LPY]) : mk(g.f(o + S_LPX), g.f(o + S_LPY));
    V2 pt0 = CACHED ? mk(cr[h + M_P0X], cr[h + M_P0Y]) : mk(g.f(o + S_PT), g.f(o + S_PT + 1));
    float minSeparation = 0.0f;
    for (int j = 0; j < 2; ++j) {
      if (j < count) {
        Xf xfA, xfB;
        xfA.q = rA_ < nbS ? rot_of(aA) : rot_identity();
        xfB.q = rB_ < nbS ? rot_of(aB) : rot_identity();
        xfA.p = cA - rmul(xfA.q, lcA);
        xfB.p = cB - rmul(xfB.q, lcB);
        V2 lpt = j == 0 ? pt0 : (CACHED ? mk(cr[h + M_P1X], cr[h + M_P1Y]) : mk(g.f(o + S_PT + 5), g.f(o + S_PT + 6)));
        V2 normal, point;
        float separation;
        if (type == MF_CIRCLES) {
          V2 pointA = xmul(xfA, lp);
          V2 pointB = xmul(xfB, pt0);
          normal = pointB - pointA;
          normalize(normal);
          point = 0.5f * (pointA + pointB);
          separation = dot(pointB - pointA, normal) - radiusA - radiusB;
        } else if (type == MF_FACE_A) {
          normal = rmul(xfA.q, ln);
          V2 planePoint = xmul(xfA, lp);
          V2 clip = xmul(xfB, lpt);
          separation = dot(clip - planePoint, normal) - radiusA - radiusB;
          point = clip;
        } else {
          normal = rmul(xfB.q, ln);
          V2 planePoint = xmul(xfB, lp);
          V2 clip = xmul(xfA, lpt);
          separation = dot(clip - planePoint, normal) - radiusA - radiusB;
          point = clip;
          normal = -normal;
        }
        V2 rA = point - cA, rB = point - cB;
        minSeparation = fminb(minSeparation, separation);
        float C = clampb(baumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
        float rnA = cross(rA, normal), rnB = cross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        V2 P = impulse * normal;
        cA -= mA * P;
        aA -= iA * cross(rA, P);
        cB += mB * P;
        aB += iB * cross(rB, P);
      }
    }
    set_hc(rA_, cA, aA);
    set_hc(rB_, cB, aB);
    return minSeparation;
  }

  // ---- revolute joints (b2RevoluteJoint) -----------------------------------------------------------------------------
  // record k (solve order) is built for joint id j; warm-start impulses were staged at record position j by load()
  // and are kept at their joint-id position: record index == joint id, solve order is a separate small table.
  BLCD_HDN void joint_init(int j, int isl, float dtRatio, float h_dt) {
    const DJoint& jd = sc.joint[j];
    const int h = kHotJoint * j;
    int rA_ = jd.a, rB_ = jd.b;
    float mA = hm(rA_), iA = hi(rA_), mB = hm(rB_), iB = hi(rB_);
    float aA = ha(rA_), aB = ha(rB_);
    V2 vA = hv(rA_), vB = hv(rB_);
    float wA = hw(rA_), wB = hw(rB_);
    Rot qA = rot_of(aA), qB = rot_of(aB);
    V2 rA = rmul(qA, jd.la - lc_of(rA_));
    V2 rB = rmul(qB, jd.lb - lc_of(rB_));
    bool fixedRotation = (iA + iB == 0.0f);
    jr[h + J_RAX] = rA.x; jr[h + J_RAY] = rA.y; jr[h + J_RBX] = rB.x; jr[h + J_RBY] = rB.y;
    jr[h + J_EXX] = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
    jr[h + J_EYX] = -rA.y * rA.x * iA - rB.y * rB.x * iB;
    jr[h + J_EZX] = -rA.y * iA - rB.y * iB;
    jr[h + J_EYY] = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
    jr[h + J_EZY] = rA.x * iA + rB.x * iB;
    jr[h + J_EZZ] = iA + iB;
    float motorMass = iA + iB;
    if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
    jr[h + J_MM] = motorMass;
    float impx = jr[h + J_IX], impy = jr[h + J_IY], impz = jr[h + J_IZ], motorImpulse = jr[h + J_MI];
    int limitState = (int)(jru(h + J_PK) & 3u);
    if (!jd.enableMotor || fixedRotation) motorImpulse = 0.0f;
    if (jd.enableLimit && !fixedRotation) {
      float jointAngle = aB - aA - jr[h + J_REF];
      if (absb(jd.upper - jd.lower) < 2.0f * kAngularSlop) limitState = 3;
      else if (jointAngle <= jd.lower) { if (limitState != 1) impz = 0.0f; limitState = 1; }
      else if (jointAngle >= jd.upper) { if (limitState != 2) impz = 0.0f; limitState = 2; }
      else { limitState = 0; impz = 0.0f; }
    } else {
      limitState = 0;
    }
    impx *= dtRatio; impy *= dtRatio; impz *= dtRatio;
    motorImpulse *= dtRatio;
    V2 P = mk(impx, impy);
    vA -= mA * P;
    wA -= iA * (cross(rA, P) + motorImpulse + impz);
    vB += mB * P;
    wB += iB * (cross(rB, P) + motorImpulse + impz);
    jr[h + J_IX] = impx; jr[h + J_IY] = impy; jr[h + J_IZ] = impz; jr[h + J_MI] = motorImpulse;
    jru(h + J_PK) = (uint32_t)limitState | ((uint32_t)isl << 4) | ((fixedRotation ? 1u : 0u) << 16);
    (void)h_dt;
    set_hv(rA_, vA, wA);
    set_hv(rB_, vB, wB);
  }

  BLCD_HD void joint_solve_velocity(int j, float h_dt) {
    const DJoint& jd = sc.joint[j];
    const int h = kHotJoint * j;
    int rA_ = jd.a, rB_ = jd.b;
    float mA = hm(rA_), iA = hi(rA_), mB = hm(rB_), iB = hi(rB_);
    V2 vA = hv(rA_), vB = hv(rB_);
    float wA = hw(rA_), wB = hw(rB_);
    uint32_t pk = jru(h + J_PK);
    int limitState = pk & 3u;
    bool fixedRotation = (pk >> 16) & 1u;
    V2 rA = mk(jr[h + J_RAX], jr[h + J_RAY]), rB = mk(jr[h + J_RBX], jr[h + J_RBY]);
    if (jd.enableMotor && limitState != 3 && !fixedRotation) {
      float Cdot = wB - wA - jr[h + J_MS];
      float impulse = -jr[h + J_MM] * Cdot;
      float oldImpulse = jr[h + J_MI];
      float maxImpulse = h_dt * jd.maxTorque;
      float ni = clampb(oldImpulse + impulse, -maxImpulse, maxImpulse);
      jr[h + J_MI] = ni;
      impulse = ni - oldImpulse;
      wA -= iA * impulse;
      wB += iB * impulse;
    }
    float exx = jr[h + J_EXX], eyx = jr[h + J_EYX], eyy = jr[h + J_EYY];
    if (jd.enableLimit && limitState != 0 && !fixedRotation) {
      float ezx = jr[h + J_EZX], ezy = jr[h + J_EZY], ezz = jr[h + J_EZZ];
      V2 Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA);
      float Cdot2 = wB - wA;
      // b2Mat33::Solve33 with ex = (exx, eyx, ezx), ey = (eyx, eyy, ezy), ez = (ezx, ezy, ezz)
      float bx = Cdot1.x, by = Cdot1.y, bz = Cdot2;
      float cx_ = eyy * ezz - ezy * ezy, cy_ = ezy * ezx - eyx * ezz, cz_ = eyx * ezy - eyy * ezx;  // cross(ey, ez)
      float det = exx * cx_ + eyx * cy_ + ezx * cz_;
      if (det != 0.0f) det = 1.0f / det;
      float ix = det * (bx * cx_ + by * cy_ + bz * cz_);
      float ux = by * ezz - bz * ezy, uy = bz * ezx - bx * ezz, uz = bx * ezy - by * ezx;        // cross(b, ez)
      float iy = det * (exx * ux + eyx * uy + ezx * uz);
      float tx = eyy * bz - ezy * by, ty = ezy * bx - eyx * bz, tz = eyx * by - eyy * bx;        // cross(ey, b)
      float iz = det * (exx * tx + eyx * ty + ezx * tz);
      ix = -ix; iy = -iy; iz = -iz;
      float jx = jr[h + J_IX], jy = jr[h + J_IY], jz = jr[h + J_IZ];
      bool reduce = false;
      if (limitState == 1) reduce = (jz + iz) < 0.0f;
      else if (limitState == 2) reduce = (jz + iz) > 0.0f;
      if (reduce) {
        V2 rhs = -Cdot1 + jz * mk(ezx, ezy);
        float d2 = exx * eyy - eyx * eyx;
        if (d2 != 0.0f) d2 = 1.0f / d2;
        float redx = d2 * (eyy * rhs.x - eyx * rhs.y), redy = d2 * (exx * rhs.y - eyx * rhs.x);
        ix = redx; iy = redy; iz = -jz;
        jx += redx; jy += redy; jz = 0.0f;
      } else {
        jx += ix; jy += iy; jz += iz;
      }
      jr[h + J_IX] = jx; jr[h + J_IY] = jy; jr[h + J_IZ] = jz;
      V2 P = mk(ix, iy);
      vA -= mA * P;
      wA -= iA * (cross(rA, P) + iz);
      vB += mB * P;
      wB += iB * (cross(rB, P) + iz);
    } else {
      V2 Cdot = vB + cross(wB, rB) - vA - cross(wA, rA);
      V2 nb_ = -Cdot;
      float d2 = exx * eyy - eyx * eyx;
      if (d2 != 0.0f) d2 = 1.0f / d2;
      V2 imp = mk(d2 * (eyy * nb_.x - eyx * nb_.y), d2 * (exx * nb_.y - eyx * nb_.x));
      jr[h + J_IX] += imp.x;
      jr[h + J_IY] += imp.y;
      vA -= mA * imp;
      wA -= iA * cross(rA, imp);
      vB += mB * imp;
      wB += iB * cross(rB, imp);
    }
    set_hv(rA_, vA, wA);
    set_hv(rB_, vB, wB);
  }

  BLCD_HD void jv_load(JV& q, int j, float h_dt) const {
    const DJoint& jd = sc.joint[j];
    const int h = kHotJoint * j;
    q.rowA = jd.a; q.rowB = jd.b;
    uint32_t pk = jru(h + J_PK);
    q.limitState = (int)(pk & 3u);
    q.flags = (jd.enableMotor ? 1 : 0) | (jd.enableLimit ? 2 : 0) | (((pk >> 16) & 1u) ? 4 : 0);
    q.rAx = jr[h + J_RAX]; q.rAy = jr[h + J_RAY]; q.rBx = jr[h + J_RBX]; q.rBy = jr[h + J_RBY];
    q.exx = jr[h + J_EXX]; q.eyx = jr[h + J_EYX]; q.ezx = jr[h + J_EZX]; q.eyy = jr[h + J_EYY]; q.ezy = jr[h + J_EZY]; q.ezz = jr[h + J_EZZ];
    q.cx = q.eyy * q.ezz - q.ezy * q.ezy; q.cy = q.ezy * q.ezx - q.eyx * q.ezz; q.cz = q.eyx * q.ezy - q.eyy * q.ezx;  // cross(ey, ez)
    float det = q.exx * q.cx + q.eyx * q.cy + q.ezx * q.cz;
    if (det != 0.0f) det = 1.0f / det;
    q.det3 = det;
    float d2 = q.exx * q.eyy - q.eyx * q.eyx;
    if (d2 != 0.0f) d2 = 1.0f / d2;
    q.det2 = d2;
    q.mm = jr[h + J_MM]; q.maxImp = h_dt * jd.maxTorque; q.ms = jr[h + J_MS];
    q.mA = hm(q.rowA); q.iA = hi(q.rowA); q.mB = hm(q.rowB); q.iB = hi(q.rowB);
    q.ix = jr[h + J_IX]; q.iy = jr[h + J_IY]; q.iz = jr[h + J_IZ]; q.mi = jr[h + J_MI];
  }

  // jv_load in two halves for the software-pipelined loop of solve_velocity: the loads of the NEXT joint's record are issued
  // before the current joint is solved and only consumed (jv_derive) after it, so their latency hides behind that arithmetic
  BLCD_HD void jv_fetch(JV& q, int j, float h_dt) const {
    const DJoint& jd = sc.joint[j];
    const int h = kHotJoint * j;
    q.rowA = jd.a; q.rowB = jd.b;
    uint32_t pk = jru(h + J_PK);
    q.limitState = (int)(pk & 3u);
    q.flags = (jd.enableMotor ? 1 : 0) | (jd.enableLimit ? 2 : 0) | (((pk >> 16) & 1u) ? 4 : 0);
    q.rAx = jr[h + J_RAX]; q.rAy = jr[h + J_RAY]; q.rBx = jr[h + J_RBX]; q.rBy = jr[h + J_RBY];
    q.mm = jr[h + J_MM]; q.maxImp = h_dt * jd.maxTorque; q.ms = jr[h + J_MS];
    q.mA = hm(q.rowA); q.iA = hi(q.rowA); q.mB = hm(q.rowB); q.iB = hi(q.rowB);
    q.ix = jr[h + J_IX]; q.iy = jr[h + J_IY]; q.iz = jr[h + J_IZ]; q.mi = jr[h + J_MI];
  }
  BLCD_HD void jv_derive(JV& q) const {
    // the effective-mass matrix is recomputed from the anchors and masses (the expressions of joint_init_velocity) rather
    // than streamed from the record: six words less traffic per joint and sweep
    q.exx = q.mA + q.mB + q.rAy * q.rAy * q.iA + q.rBy * q.rBy * q.iB;
    q.eyx = -q.rAy * q.rAx * q.iA - q.rBy * q.rBx * q.iB;
    q.ezx = -q.rAy * q.iA - q.rBy * q.iB;
    q.eyy = q.mA + q.mB + q.rAx * q.rAx * q.iA + q.rBx * q.rBx * q.iB;
    q.ezy = q.rAx * q.iA + q.rBx * q.iB;
    q.ezz = q.iA + q.iB;
    q.cx = q.eyy * q.ezz - q.ezy * q.ezy; q.cy = q.ezy * q.ezx - q.eyx * q.ezz; q.cz = q.eyx * q.ezy - q.eyy * q.ezx;  // cross(ey, ez)
    float det = q.exx * q.cx + q.eyx * q.cy + q.ezx * q.cz;
    if (det != 0.0f) det = 1.0f / det;
    q.det3 = det;
    float d2 = q.exx * q.eyy - q.eyx * q.eyx;
    if (d2 != 0.0f) d2 = 1.0f / d2;
    q.det2 = d2;
  }

  BLCD_HD void jv_save(const JV& q, int j) {
    const int h = kHotJoint * j;
    jr[h + J_IX] = q.ix; jr[h + J_IY] = q.iy; jr[h + J_IZ] = q.iz; jr[h + J_MI] = q.mi;
  }

  // b2RevoluteJoint::SolveVelocityConstraints on a register-resident record, written WITHOUT branches: in a warp of 32
  // worlds nearly every iteration has some lane in each of Box2D's cases (motor on/off, limit inactive / active /
  // clamped), so the warp would walk all branch bodies anyway; selects keep the instruction stream straight (no fetch
  // redirects) and the arithmetic of the selected case is exactly Box2D's (x + 0 == x, x - m * 0 == x).
  BLCD_HD void jv_solve(JV& q) const {
    V2 vA = hv(q.rowA), vB = hv(q.rowB);
    float wA = hw(q.rowA), wB = hw(q.rowB);
    const bool rot = (q.flags & 4) == 0;  // !fixedRotation
    const V2 rA = mk(q.rAx, q.rAy), rB = mk(q.rBx, q.rBy);
    {  // motor
      const bool on = (q.flags & 1) && q.limitState != 3 && rot;
      float Cdot = wB - wA - q.ms;
      float impulse = -q.mm * Cdot;
      float oldImpulse = q.mi;
      float ni = clampb(oldImpulse + impulse, -q.maxImp, q.maxImp);
      ni = on ? ni : oldImpulse;
      impulse = ni - oldImpulse;
      q.mi = ni;
      wA -= q.iA * impulse;
      wB += q.iB * impulse;
    }
    V2 Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA);
    const bool limited = (q.flags & 2) && q.limitState != 0 && rot;
    // 3x3 solve (limit active)
    float bx = Cdot1.x, by = Cdot1.y, bz = wB - wA;
    float i3x = q.det3 * (bx * q.cx + by * q.cy + bz * q.cz);
    float ux = by * q.ezz - bz * q.ezy, uy = bz * q.ezx - bx * q.ezz, uz = bx * q.ezy - by * q.ezx;        // cross(b, ez)
    float i3y = q.det3 * (q.exx * ux + q.eyx * uy + q.ezx * uz);
    float tx = q.eyy * bz - q.ezy * by, ty = q.ezy * bx - q.eyx * bz, tz = q.eyx * by - q.eyy * bx;        // cross(ey, b)
    float i3z = q.det3 * (q.exx * tx + q.eyx * ty + q.ezx * tz);
    i3x = -i3x; i3y = -i3y; i3z = -i3z;
    float sum = q.iz + i3z;
    const bool reduce = limited && ((q.limitState == 1 && sum < 0.0f) || (q.limitState == 2 && sum > 0.0f));
    // 2x2 solve: point-to-point (rhs = -Cdot1) or the clamped limit case (rhs = -Cdot1 + impulse.z * ez.xy)
    V2 add = q.iz * mk(q.ezx, q.ezy);
    add.x = reduce ? add.x : 0.0f;
    add.y = reduce ? add.y : 0.0f;
    V2 rhs = -Cdot1 + add;
    float i2x = q.det2 * (q.eyy * rhs.x - q.eyx * rhs.y), i2y = q.det2 * (q.exx * rhs.y - q.eyx * rhs.x);
    const bool use3 = limited && !reduce;
    float ix = use3 ? i3x : i2x, iy = use3 ? i3y : i2y;
    float iz = use3 ? i3z : (limited ? -q.iz : 0.0f);
    q.ix += ix;
    q.iy += iy;
    q.iz = use3 ? sum : (limited ? 0.0f : q.iz);
    V2 P = mk(ix, iy);
    vA -= q.mA * P;
    wA -= q.iA * (cross(rA, P) + iz);
    vB += q.mB * P;
    wB += q.iB * (cross(rB, P) + iz);
    set_hv(q.rowA, vA, wA);
    set_hv(q.rowB, vB, wB);
  }

  BLCD_HD bool joint_solve_position(int j) {
    const DJoint& jd = sc.joint[j];
    const int h = kHotJoint * j;
    int rA_ = jd.a, rB_ = jd.b;
    float mA = hm(rA_), iA = hi(rA_), mB = hm(rB_), iB = hi(rB_);
    V2 cA = hc(rA_), cB = hc(rB_);
    float aA = ha(rA_), aB = ha(rB_);
    uint32_t pk = jru(h + J_PK);
    int limitState = pk & 3u;
    bool fixedRotation = (pk >> 16) & 1u;
    float angularError = 0.0f, positionError = 0.0f;
    {
      // angular limit, branch-free: the three active cases of b2RevoluteJoint::SolvePositionConstraints differ only in
      // the reference angle, the slop shift and the clamp interval; an inactive limit applies a zero impulse
      const bool limited = jd.enableLimit && limitState != 0 && !fixedRotation;
      float angle = aB - aA - jr[h + J_REF];
      float motorMass = jr[h + J_MM];
      float C0 = angle - (limitState == 2 ? jd.upper : jd.lower);
      float shift = limitState == 1 ? kAngularSlop : (limitState == 2 ? -kAngularSlop : 0.0f);
      float lo = limitState == 2 ? 0.0f : -kMaxAngularCorrection;
      float hi = limitState == 1 ? 0.0f : kMaxAngularCorrection;
      float C = clampb(C0 + shift, lo, hi);
      float err = limitState == 3 ? absb(C) : (limitState == 1 ? -C0 : C0);
      float limitImpulse = limited ? -motorMass * C : 0.0f;
      angularError = limited ? err : 0.0f;
      aA -= iA * limitImpulse;
      aB += iB * limitImpulse;
    }
    {
      Rot qA = rot_of(aA), qB = rot_of(aB);
      V2 rA = rmul(qA, jd.la - lc_of(rA_));
      V2 rB = rmul(qB, jd.lb - lc_of(rB_));
      V2 C = cB + rB - cA - rA;
      positionError = len(C);
      float kxx = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
      float kxy = -iA * rA.x * rA.y - iB * rB.x * rB.y;
      float kyy = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
      float det = kxx * kyy - kxy * kxy;
      if (det != 0.0f) det = 1.0f / det;
      V2 imp = -mk(det * (kyy * C.x - kxy * C.y), det * (kxx * C.y - kxy * C.x));
      cA -= mA * imp;
      aA -= iA * cross(rA, imp);
      cB += mB * imp;
      aB += iB * cross(rB, imp);
    }
    set_hc(rA_, cA, aA);
    set_hc(rB_, cB, aB);
    return positionError <= kLinearSlop && angularError <= kAngularSlop;
  }

  // ---- b2World::Solve -------------------------------------------------------------------------------------------------
  // Split into the five phases a sub-step's solve consists of.  The fused kernels (k_step / k_rollout) run them back to
  // back from one thread; the phase pipeline (blcd_pipeline.cuh) runs each as its own kernel with the records in between
  // spilled to a per-world scratch area in HBM -- same functions, same order, same results.
  BLCD_HD void solve(float h_dt, float dtRatio) {
    solve_setup(h_dt, dtRatio);
    if (sc.align_mode >= 1) phase_align(0); else reconverge();   // end of: narrow phase, islands, constraint setup
    solve_velocity(h_dt);
    solve_integrate(h_dt);
    if (sc.align_mode >= 2) phase_align(1); else reconverge();   // end of: velocity iterations
    solve_position();
    if (sc.align_mode >= 4) phase_align(2); else reconverge();   // end of: position iterations
    solve_finish(h_dt);
  }

  // islands (Box2D's DFS order), velocity integration, contact / joint velocity-constraint records, warm starting
  BLCD_HD void solve_setup(float h_dt, float dtRatio) {
    const int nb = sc.nb;
    uint8_t stack[kMaxBodies];
    PairMask cflag = PairMask::none();
    uint32_t jflag = 0u;
    for (int b = 0; b < kMaxBodies; ++b) islandOf[b] = -1;
    nIslands = 0; nc = 0; njo = 0;
    stage_rows();
    // island DFS in Box2D's order: seeds newest body first; per body its contact edges (newest first) then joint edges
    for (int seed = nb - 1; seed >= 0; --seed) {
      if (islandOf[seed] >= 0 || !is_awake(seed)) continue;
      int isl = nIslands++;
      int sp = 0;
      stack[sp++] = (uint8_t)seed;
      islandOf[seed] = (int8_t)isl;
      while (sp > 0) {
        int b = stack[--sp];
        set_awake(b);
        int fb_ = sc.nw + b;
        for (int k = 0; k < ncl; ++k) {
          int p = clist[k];
          int fa = sc.pair[p].fa, fb = sc.pair[p].fb;
          if (fa != fb_ && fb != fb_) continue;
          if (cflag.test(p)) continue;
          int s = pslot[p];
          if (s < 0) continue;  // not touching
          cflag.set(p);
          if (nc < sc.maxm) {
            // record filled later (needs integrated velocities); remember slot + island in the packed word
            cru(kHotCon * nc + C_PK) = ((uint32_t)s << 12) | ((uint32_t)isl << 20);
            ++nc;
          }
          int other = fa == fb_ ? fb : fa;
          if (other < sc.nw) continue;  // static bodies are not expanded
          int ob = other - sc.nw;
          if (islandOf[ob] >= 0) continue;
          stack[sp++] = (uint8_t)ob;
          islandOf[ob] = (int8_t)isl;
        }
        const DBody& bd = sc.body[b];
        for (int e = 0; e < bd.njedge; ++e) {
          int j = bd.jedge[e];
          if ((jflag >> j) & 1u) continue;
          jflag |= 1u << j;
          jorder[njo] = (uint8_t)j; jisl[njo] = (uint8_t)isl;
          ++njo;
          int ob = sc.joint[j].a == b ? sc.joint[j].b : sc.joint[j].a;
          if (islandOf[ob] >= 0) continue;
          stack[sp++] = (uint8_t)ob;
          islandOf[ob] = (int8_t)isl;
        }
      }
    }
    cnt[BLCD_CNT_CONTACTS] += (uint32_t)nc;
    // b2Island::Solve: integrate velocities (gravity, damping)
    for (int b = 0; b < nb; ++b) {
      if (islandOf[b] < 0) continue;
      c0[b] = c[b]; a0[b] = a[b];
      V2 vv = v[b];
      float ww = w[b];
      float im = hm(b), ii = hi(b);
      vv += h_dt * (1.0f * sc.gravity + im * mk(0.0f, 0.0f));
      ww += h_dt * ii * 0.0f;
      if (sc.flags & BLCD_FLAG_DAMPING_2_3_0) {
        vv *= clampb(1.0f - h_dt * sc.body[b].linDamp, 0.0f, 1.0f);
        ww *= clampb(1.0f - h_dt * sc.body[b].angDamp, 0.0f, 1.0f);
      } else {
        vv *= 1.0f / (1.0f + h_dt * sc.body[b].linDamp);
        ww *= 1.0f / (1.0f + h_dt * sc.body[b].angDamp);
      }
      set_hv(b, vv, ww);
    }
    for (int k = 0; k < nc; ++k) {
      uint32_t pk = cru(kHotCon * k + C_PK);
      contact_init(k, (int)((pk >> 12) & 255u), (int)(pk >> 20), dtRatio, true);
    }
    for (int k = 0; k < nc; ++k) contact_warm_start(k);
    for (int k = 0; k < njo; ++k) joint_init(jorder[k], jisl[k], dtRatio, h_dt);
  }

  // b2Island::Solve's velocity iterations + b2ContactSolver::StoreImpulses
  // STREAM_CONTACTS: the pipeline's velocity kernel (128 registers, nothing else live) also streams the contact records of
  // many-joint scenes through registers; the fused kernels do not (no registers to spare: SpiderCube -6 % when they did)
  template <bool STREAM_CONTACTS = false>
  BLCD_HD void solve_velocity(float h_dt) {
    const int vi = sc.vel_iters;
    if (njo <= 3) {
      // up to three joints (every reference robot in scope) live in registers for the whole loop; contacts, whose
      // number is data dependent, stay in shared memory
      JV ja, jb, jc;
      if (njo > 0) jv_load(ja, jorder[0], h_dt);
      if (njo > 1) jv_load(jb, jorder[1], h_dt);
      if (njo > 2) jv_load(jc, jorder[2], h_dt);
      const int njo_ = njo, nc_ = nc;
      for (int it = 0; it < vi; ++it) {
        if (njo_ > 0) jv_solve(ja);
        if (njo_ > 1) jv_solve(jb);
        if (njo_ > 2) jv_solve(jc);
        bool changed = false;
        for (int k = 0; k < nc_; ++k) changed |= contact_solve_velocity(k);
        // a sweep that applied no impulse leaves the state untouched, so every later sweep repeats it exactly
        if (njo_ == 0 && !changed) break;
      }
      if (njo > 0) jv_save(ja, jorder[0]);
      if (njo > 1) jv_save(jb, jorder[1]);
      if (njo > 2) jv_save(jc, jorder[2]);
    } else {
      // more joints than the register file holds (the crab-class robots have 16): their records stream from thread-local
      // memory every sweep, the next one being fetched while the current one is solved
      const int njo_ = njo, nc_ = nc;
#if BLCD_PROFILE_ID == 1
      // The contact records stream the same way, through two register copies (ra: being solved, rb: arriving).  Large-scene
      // build only: in the small build (4-7 joints: none of the reference's robots) the two copies land on the stack of the
      // fused kernel and cost the THREE-joint robots 12 % (Urchin, 32 768 worlds: 14.5 -> 12.7 M env-steps/s, measured).
      if (STREAM_CONTACTS) {
        JV cur, nxt;
        float ra[kHotCon], rb[kHotCon];
        jv_fetch(cur, jorder[0], h_dt);
        for (int it = 0; it < vi; ++it) {
          jv_derive(cur);
          for (int k = 0; k < njo_; ++k) {
            const int j = jorder[k];
            if (k + 1 < njo_) jv_fetch(nxt, jorder[k + 1], h_dt);
            else if (nc_ > 0) rec_fetch(ra, 0);
            jv_solve(cur);
            jv_save(cur, j);
            if (k + 1 < njo_) { jv_derive(nxt); cur = nxt; }
          }
          for (int k = 0; k < nc_; ++k) {
            if (k + 1 < nc_) rec_fetch(rb, k + 1);
            else if (it + 1 < vi) jv_fetch(cur, jorder[0], h_dt);
            contact_solve_velocity_rec(ra);
            rec_save(ra, k);
            if (k + 1 < nc_) {
#pragma unroll
              for (int i = 0; i < kHotCon; ++i) ra[i] = rb[i];
            }
          }
          if (nc_ == 0 && it + 1 < vi) jv_fetch(cur, jorder[0], h_dt);
        }
      } else
#endif
      for (int it = 0; it < vi; ++it) {
        JV cur, nxt;
        jv_fetch(cur, jorder[0], h_dt);
        jv_derive(cur);
        for (int k = 0; k < njo_; ++k) {
          const int j = jorder[k];
          if (k + 1 < njo_) jv_fetch(nxt, jorder[k + 1], h_dt);
          jv_solve(cur);
          jv_save(cur, j);
          if (k + 1 < njo_) { jv_derive(nxt); cur = nxt; }
        }
        for (int k = 0; k < nc_; ++k) contact_solve_velocity(k);
      }
    }
    for (int k = 0; k < nc; ++k) contact_store_impulses(k);
  }

  // b2Island::Solve: integrate positions (maxTranslation / maxRotation clamps)
  BLCD_HD void solve_integrate(float h_dt) {
    const int nb = sc.nb;
    for (int b = 0; b < nb; ++b) {
      if (islandOf[b] < 0) continue;
      V2 cc = hc(b), vv = hv(b);
      float aa = ha(b), ww = hw(b);
      V2 translation = h_dt * vv;
      if (dot(translation, translation) > kMaxTranslationSq) { float ratio = kMaxTranslation / len(translation); vv *= ratio; }
      float rotation = h_dt * ww;
      if (rotation * rotation > kMaxRotationSq) { float ratio = kMaxRotation / absb(rotation); ww *= ratio; }
      cc += h_dt * vv;
      aa += h_dt * ww;
      set_hc(b, cc, aa);
      set_hv(b, vv, ww);
    }
  }

  // the same with the positions read from and written to the scratch rows (lean velocity kernel), velocities spilled too
  BLCD_HD void solve_integrate_x(float h_dt) {
    const int nb = sc.nb;
    for (int b = 0; b < nb; ++b) {
      V2 vv = hv(b);
      float ww = hw(b);
      if (islandOf[b] >= 0) {
        const int o = sc.x_rows + 8 * b;
        V2 cc = mk(x.f(o + 3), x.f(o + 4));
        float aa = x.f(o + 5);
        V2 translation = h_dt * vv;
        if (dot(translation, translation) > kMaxTranslationSq) { float ratio = kMaxTranslation / len(translation); vv *= ratio; }
        float rotation = h_dt * ww;
        if (rotation * rotation > kMaxRotationSq) { float ratio = kMaxRotation / absb(rotation); ww *= ratio; }
        cc += h_dt * vv;
        aa += h_dt * ww;
        x.f(o + 3) = cc.x; x.f(o + 4) = cc.y; x.f(o + 5) = aa;
      }
      x.f(sc.x_rows + 8 * b) = vv.x; x.f(sc.x_rows + 8 * b + 1) = vv.y; x.f(sc.x_rows + 8 * b + 2) = ww;
    }
  }

  // position iterations; every island stops on its own convergence
  BLCD_HD void solve_position() {
    islDone = 0u;
    // the manifolds are copied out of their HBM slots once (into the point-record words of the contact records, which
    // the velocity phase no longer needs: its impulses are already stored) instead of being re-read every sweep
    for (int k = 0; k < nc; ++k) pos_cache_manifold(k);
    for (int it = 0; it < sc.pos_iters; ++it)
      if (solve_position_sweep<true>()) break;
  }

  // one position iteration over the islands that have not converged yet; true when none is left.  (The pipeline's position
  // kernel calls this once per loop trip for whatever world a lane currently holds, blcd_pipeline.cuh.)
  template <bool CACHED = false>
  BLCD_HD bool solve_position_sweep() {
    const uint32_t islAll = (1u << nIslands) - 1u;
    if (islDone == islAll) return true;
    uint32_t bad = 0u;
    cnt[BLCD_CNT_POS_ITERS] += (uint32_t)(nIslands - popc(islDone));
    for (int k = 0; k < nc; ++k) {
      int isl = (int)(cru(kHotCon * k + C_PK) >> 20);
      if ((islDone >> isl) & 1u) continue;
      float ms = contact_solve_position_t<CACHED>(k, kBaumgarte);
      if (!(ms >= -3.0f * kLinearSlop)) bad |= 1u << isl;
    }
    for (int k = 0; k < njo; ++k) {
      int isl = jisl[k];
      if ((islDone >> isl) & 1u) continue;
      if (!joint_solve_position(jorder[k])) bad |= 1u << isl;
    }
    islDone |= ~bad & islAll;
    return islDone == islAll;
  }

  // write back + SynchronizeTransform, sleeping, broad phase
  BLCD_HD void solve_finish(float h_dt) {
    const int nb = sc.nb;
    Xf xf1[kMaxBodies];
    for (int b = 0; b < nb; ++b) {
      if (islandOf[b] < 0) continue;
      xf1[b] = xf[b];  // transform at (c0, a0)
      c[b] = hc(b); a[b] = ha(b);
      v[b] = hv(b); w[b] = hw(b);
      xf[b] = xf_of(c[b], a[b], lc_of(b));
    }
    // sleeping (per island)
    if (!(sc.flags & BLCD_FLAG_NO_SLEEP)) {
      for (int isl = 0; isl < nIslands; ++isl) {
        float minSleep = kMaxFloat;
        for (int b = 0; b < nb; ++b) {
          if (islandOf[b] != isl) continue;
          if (w[b] * w[b] > kAngSleepTol * kAngSleepTol || dot(v[b], v[b]) > kLinSleepTol * kLinSleepTol) {
            sleepT[b] = 0.0f;
            minSleep = 0.0f;
          } else {
            sleepT[b] += h_dt;
            minSleep = fminb(minSleep, sleepT[b]);
          }
        }
        if (minSleep >= kTimeToSleep && ((islDone >> isl) & 1u)) {
          for (int b = 0; b < nb; ++b) {
            if (islandOf[b] != isl) continue;
            awake &= ~(1u << b);
            sleepT[b] = 0.0f;
            v[b] = mk(0.0f, 0.0f);
            w[b] = 0.0f;
          }
        }
      }
    }
    // broad phase: newest body first, then b2ContactManager::FindNewContacts
    moved = 0u;
    for (int b = nb - 1; b >= 0; --b) {
      if (islandOf[b] < 0) continue;
      sync_fixture(b, xf1[b]);
    }
    find_new_contacts(moved);
    moved = 0u;
  }

  BLCD_HD static int popc(uint32_t x) {
    int n = 0;
    while (x) { x &= x - 1u; ++n; }
    return n;
  }

  // ---- b2World::SolveTOI: continuous collision of awake dynamic bodies against the static walls ----------------------
  BLCD_HDN void solve_toi(float h_dt) {
    const int nb = sc.nb, nw = sc.nw;
    float toi[kMaxPairs];
    uint8_t toiCount[kMaxPairs];
    PairMask toiValid = PairMask::none(), enabled = PairMask::all();
    for (int b = 0; b < nb; ++b) alpha0[b] = 0.0f;
    for (int i = 0; i < BLCD_MAX_WALLS; ++i) walpha0[i] = 0.0f;
    for (int k = 0; k < ncl; ++k) { toiCount[clist[k]] = 0; toi[clist[k]] = 1.0f; }
    uint32_t toiMask[kMaxBodies];
    for (int b = 0; b < kMaxBodies; ++b)
      if (b < nb) toiMask[b] = is_awake(b) ? toi_prefilter(b) : 0xFFFFFFFFu;
    for (;;) {
      // Pass 1, in contact-list order: bookkeeping of b2World::SolveTOI's scan (sweeps onto a common interval, counters)
      // and collection of the pairs that really need a time-of-impact query.  The queries are pure functions of the
      // sweeps, so they are deferred to pass 2, where all lanes of the warp run theirs together instead of each lane
      // hitting the long GJK / root-finder code at a different list position.  If a body's sweep is about to be advanced
      // while queries on it are pending, those are flushed first, so every query still sees exactly Box2D's inputs.
      uint8_t pend[16];
      float pend_al0[16];   // common interval start of each pending query, as of its place in the list
      int npend = 0;
      for (int k = 0; k < ncl; ++k) {
        int p = clist[k];
        if (!enabled.test(p)) continue;
        if (toiCount[p] > kMaxSubSteps) continue;
        if (toiValid.test(p)) continue;
        int fa = sc.pair[p].fa, fb = sc.pair[p].fb;
        if (fa >= nw) continue;          // two non-bullet dynamic bodies
        int b = fb - nw;
        if (!is_awake(b)) continue;      // neither side active
        // put both sweeps on the same interval
        if (walpha0[fa] < alpha0[b]) walpha0[fa] = alpha0[b];
        else if (alpha0[b] < walpha0[fa]) {
          toi_flush(pend, pend_al0, npend, toi, toiValid);
          Sweep sb = body_sweep(b);
          sweep_advance(sb, walpha0[fa]);
          c0[b] = sb.c0; a0[b] = sb.a0; alpha0[b] = sb.alpha0;
          toiMask[b] = toi_prefilter(b);   // the sweep's start moved
        }
        ++cnt[BLCD_CNT_TOI_CALLS];
        if (!((toiMask[b] >> fa) & 1u)) {
          toi[p] = 1.0f;   // b2TimeOfImpact would come back e_separated / e_failed: alpha = 1, no side effects
          toiValid.set(p);
        } else {
          if (npend == 16) toi_flush(pend, pend_al0, npend, toi, toiValid);
          pend_al0[npend] = walpha0[fa];
          pend[npend++] = (uint8_t)p;
        }
      }
      // Pass 2: the deferred queries
      toi_flush(pend, pend_al0, npend, toi, toiValid);
      // Pass 3: earliest time of impact, first in list order on ties
      int minPair = -1;
      float minAlpha = 1.0f;
      for (int k = 0; k < ncl; ++k) {
        int p = clist[k];
        if (!enabled.test(p)) continue;
        if (toiCount[p] > kMaxSubSteps) continue;
        if (!toiValid.test(p)) continue;
        float alpha = toi[p];
        if (alpha < minAlpha) { minPair = p; minAlpha = alpha; }
      }
      ph(4);   // (diagnostic build) candidate scan: skip tests + time-of-impact queries
      if (minPair < 0 || 1.0f - 10.0f * kEps < minAlpha) break;
      const int wl = sc.pair[minPair].fa, b = sc.pair[minPair].fb - nw;
      // advance both bodies to the time of impact
      V2 bk_c0 = c0[b], bk_c = c[b];
      float bk_a0 = a0[b], bk_a = a[b], bk_alpha0 = alpha0[b], bk_walpha = walpha0[wl];
      {
        Sweep sb = body_sweep(b);
        sweep_advance(sb, minAlpha);
        c0[b] = sb.c0; a0[b] = sb.a0; alpha0[b] = sb.alpha0;
        c[b] = sb.c0; a[b] = sb.a0;
        xf[b] = xf_of(c[b], a[b], lc_of(b));
        walpha0[wl] = minAlpha;
      }
      bool touching = update_contact(minPair);
      toiValid.clear(minPair);
      ++toiCount[minPair];
      if (!touching) {
        enabled.clear(minPair);
        c0[b] = bk_c0; c[b] = bk_c; a0[b] = bk_a0; a[b] = bk_a; alpha0[b] = bk_alpha0; walpha0[wl] = bk_walpha;
        xf[b] = xf_of(c[b], a[b], lc_of(b));
        continue;
      }
      ++cnt[BLCD_CNT_TOI_EVENTS];
      set_awake(b);
      // mini island: body b, the wall, plus the other walls b touches at this pose
      int ntc = 0;
      uint32_t wallIn = 1u << wl;
      PairMask inIsland = PairMask::bit(minPair);
      cru(kHotCon * ntc + C_PK) = ((uint32_t)pslot[minPair] << 12);
      ++ntc;
      for (int k = 0; k < ncl; ++k) {
        int p = clist[k];
        int fa = sc.pair[p].fa, fb = sc.pair[p].fb;
        if (fb != nw + b && fa != nw + b) continue;
        if (ntc == sc.maxm) break;
        if (inIsland.test(p)) continue;
        if (fa >= nw) continue;  // other body dynamic, no bullets: skipped
        float bkw = walpha0[fa];
        if (!((wallIn >> fa) & 1u)) walpha0[fa] = minAlpha;
        bool t2 = update_contact(p);
        enabled.set(p);     // b2Contact::Update re-enables
        if (!t2) { walpha0[fa] = bkw; continue; }
        inIsland.set(p);
        cru(kHotCon * ntc + C_PK) = ((uint32_t)pslot[p] << 12);
        ++ntc;
        wallIn |= 1u << fa;
      }
      enabled.set(minPair);
      // b2Island::SolveTOI
      float subDt = (1.0f - minAlpha) * h_dt;
      stage_rows();
      for (int k = 0; k < ntc; ++k) {  // position constraints only need the packed rows + slot
        uint32_t pk = cru(kHotCon * k + C_PK);
        int s = (int)((pk >> 12) & 255u);
        int p = (int)(g.u(slot_base(s) + S_HDR) & 0xFFu);
        int fA, fB;
        pair_ab(p, &fA, &fB);
        cru(kHotCon * k + C_PK) = (uint32_t)row_of(fA) | ((uint32_t)row_of(fB) << 5) | ((uint32_t)s << 12);
      }
      for (int it = 0; it < 20; ++it) {
        float minSep = 0.0f;
        for (int k = 0; k < ntc; ++k) minSep = fminb(minSep, contact_solve_position(k, kToiBaumgarte));
        if (minSep >= -1.5f * kLinearSlop) break;
      }
      c0[b] = hc(b); a0[b] = ha(b);  // leap of faith to the new safe state
      for (int k = 0; k < ntc; ++k) {
        uint32_t pk = cru(kHotCon * k + C_PK);
        contact_init(k, (int)((pk >> 12) & 255u), 0, 1.0f, false);
      }
      for (int it = 0; it < sc.vel_iters; ++it) {
        bool changed = false;
        for (int k = 0; k < ntc; ++k) changed |= contact_solve_velocity(k);
        if (!changed) break;  // fixed point reached bit-for-bit: the remaining sweeps would all be no-ops
      }
      {
        V2 cc = hc(b), vv = hv(b);
        float aa = ha(b), ww = hw(b);
        V2 translation = subDt * vv;
        if (dot(translation, translation) > kMaxTranslationSq) { float ratio = kMaxTranslation / len(translation); vv *= ratio; }
        float rotation = subDt * ww;
        if (rotation * rotation > kMaxRotationSq) { float ratio = kMaxRotation / absb(rotation); ww *= ratio; }
        cc += subDt * vv;
        aa += subDt * ww;
        c[b] = cc; a[b] = aa; v[b] = vv; w[b] = ww;
        xf[b] = xf_of(c[b], a[b], lc_of(b));
      }
      // synchronize the broad phase, invalidate the cached TOIs of everything touching the displaced body
      moved = 0u;
      sync_fixture(b, xf_of(c0[b], a0[b], lc_of(b)));
      for (int k = 0; k < ncl; ++k) {
        int p = clist[k];
        if (sc.pair[p].fb == nw + b || sc.pair[p].fa == nw + b) toiValid.clear(p);
      }
      int ncl0 = ncl;
      toiMask[b] = toi_prefilter(b);   // new sweep after the event
      ph(3);   // (diagnostic build) TOI event: advance, mini-island solve
      find_new_contacts(moved);
      for (int k = 0; k < ncl - ncl0; ++k) { toiCount[clist[k]] = 0; toi[clist[k]] = 1.0f; toiValid.clear(clist[k]); enabled.set(clist[k]); }
      moved = 0u;
    }
  }

  // Exact pre-filter for the TOI queries of body b: bit k of the result is set if b2TimeOfImpact against wall k could
  // report e_touching.  It can only do so if, at some time of the sweep, the two cores (polygon without its skin / circle
  // centre, wall segment) come closer than target + tolerance.  Every core point moves as c(t) + R(theta(t)) r with c and
  // theta linear in t, so its signed distance to the wall's line stays above the smaller of its two end values minus the
  // sag of the rotation arc, |r| dtheta^2 / 8.  If that lower bound (with 1 mm of margin for rounding) clears the
  // threshold for every vertex, the query's answer -- alpha = 1 -- is known without running GJK and the root finder.
  // Evaluated once per body (all lanes together) instead of once per candidate pair.
  BLCD_HDN uint32_t toi_prefilter(int b) const {
    const DShape& S = bshape(b);
    Xf x0 = xf_of(c0[b], a0[b], lc_of(b));
    const Xf& x1 = xf[b];
    float da = a[b] - a0[b];
    float sag = S.rmax * da * da * 0.125f * 1.01f;
    V2 p0[BLCD_MAX_VERTS], p1[BLCD_MAX_VERTS];
    for (int i = 0; i < BLCD_MAX_VERTS; ++i)
      if (i < S.count) { p0[i] = xmul(x0, S.v[i]); p1[i] = xmul(x1, S.v[i]); }
    uint32_t mask = 0u;
    for (int k = 0; k < sc.nw; ++k) {
      V2 n = sc.wall_n[k];
      float d0 = sc.wall_d[k];
      float totalRadius = sc.wall[k].radius + S.radius;
      float thresh = fmaxb(kLinearSlop, totalRadius - 3.0f * kLinearSlop) + 0.25f * kLinearSlop + 1.0e-3f;
      float lo = kMaxFloat, hi = -kMaxFloat;
      for (int i = 0; i < BLCD_MAX_VERTS; ++i) {
        if (i < S.count) {
          float s0 = dot(n, p0[i]) - d0, s1 = dot(n, p1[i]) - d0;
          lo = fminb(lo, fminb(s0, s1));
          hi = fmaxb(hi, fmaxb(s0, s1));
        }
      }
      bool clear = (lo - sag > thresh) || (-hi - sag > thresh);   // one side of the line for the whole sweep, far enough
      if (!clear || !(dot(n, n) > 0.5f)) mask |= 1u << k;
    }
    return mask;
  }

  // run the pending time-of-impact queries (pairs wall fa x body b) against the current sweeps
  BLCD_HD void toi_flush(const uint8_t* pend, const float* pend_al0, int& npend, float* toi, PairMask& toiValid) {
    for (int j = 0; j < npend; ++j) {
      int p = pend[j];
      int fa = sc.pair[p].fa, b = sc.pair[p].fb - sc.nw;
      float al0 = pend_al0[j];   // == alpha0[b]: a body's own sweep is never advanced while it has queries pending
      Sweep sA;
      sA.lc = mk(0.0f, 0.0f); sA.c0 = mk(0.0f, 0.0f); sA.c = mk(0.0f, 0.0f); sA.a0 = 0.0f; sA.a = 0.0f; sA.alpha0 = al0;
      float t;
      int state = time_of_impact(&t, sc.wall[fa], sA, bshape(b), body_sweep(b), true);
      toi[p] = state == TOI_TOUCHING ? fminb(al0 + (1.0f - al0) * t, 1.0f) : 1.0f;
      toiValid.set(p);
    }
    npend = 0;
  }

  BLCD_HD Sweep body_sweep(int b) const {
    Sweep s;
    s.lc = lc_of(b); s.c0 = c0[b]; s.c = c[b]; s.a0 = a0[b]; s.a = a[b]; s.alpha0 = alpha0[b];
    return s;
  }

  // ---- b2World::Step -------------------------------------------------------------------------------------------------
  BLCD_HD void b2_step() {
    const float dt = sc.dt;
    if (sc.align_mode >= 3) phase_align(4); else reconverge();   // end of: TOI of the previous sub-step (+ observation)
    substep_collide();
    solve(dt, inv_dt0 * dt);
    if (sc.align_mode >= 3) phase_align(3); else reconverge();   // end of: write-back, sleep, broad phase
    if (!(sc.flags & BLCD_FLAG_NO_TOI)) solve_toi(dt);
    reconverge();
    substep_end();
  }

  // b2World::Step up to b2World::Solve: pending FindNewContacts, then b2ContactManager::Collide
  BLCD_HD void substep_collide() {
    if (newFixture) {
      find_new_contacts((1u << (sc.nw + sc.nb)) - 1u);
      newFixture = false;
    }
    collide();
  }

  // end of b2World::Step: m_inv_dt0, plus this library's diagnostic counters
  BLCD_HD void substep_end() {
    inv_dt0 = 1.0f / sc.dt;
    ++cnt[BLCD_CNT_SUBSTEPS];
    for (int s = 0; s < sc.maxm; ++s)
      if ((slotUsed >> s) & 1u) cnt[BLCD_CNT_MANIFOLD_POINTS] += (g.u(slot_base(s) + S_HDR) >> 16) & 0xFFu;
    cnt[BLCD_CNT_SLEEP_STEPS] += (uint32_t)(sc.nb - popc(awake & ((1u << sc.nb) - 1u)));
  }

  // WorldEnv.step (world_env.py:431-445): motor speeds from the clipped action (SetMotorSpeed wakes both bodies)
  BLCD_HD void env_step_begin(const float* action) {
    ep_t += 1;
    for (int j = 0; j < sc.nj; ++j) {
      const DJoint& jd = sc.joint[j];
      if (jd.act < 0) continue;
      double av = (double)action[jd.act];
      av = av < -1.0 ? -1.0 : (av > 1.0 ? 1.0 : av);
      set_awake(jd.a);
      set_awake(jd.b);
      jr[kHotJoint * j + J_MS] = (float)(jd.speed * av);
    }
  }

  // WorldEnv.step (world_env.py:431-452): motor speeds from the clipped action, then n_substeps world steps
  BLCD_HD void env_step(const float* action) {
    if (sc.align_mode >= 5) phase_align(4);
    env_step_begin(action);
    for (int s = 0; s < sc.nsub; ++s) b2_step();
  }

  // Would b2World::SolveTOI do anything for this world?  Its first scan visits, in contact-list order, every candidate pair
  // of an awake body with a wall, counts it, and queries the time of impact unless the pre-filter already knows the answer
  // (alpha = 1).  If every pair is pre-filtered the scan finds no event and SolveTOI changes nothing but that counter, so
  // the pipeline finishes the sub-step right away and only worlds with a real query go on to the TOI kernel.
  // Call with c0 / a0 = the pre-solve pose and c / a / xf = the solved one.  Returns true if a query is needed; otherwise
  // the counter has been advanced exactly as SolveTOI would have.
  BLCD_HD bool toi_needed() {
    const int nw = sc.nw;
    int eligible = 0;
    uint32_t bodies = 0u;    // awake bodies with a wall pair in the list
    PairMask seen = PairMask::none();
    for (int k = 0; k < ncl; ++k) {
      int p = clist[k];
      int fa = sc.pair[p].fa, fb = sc.pair[p].fb;
      if (fa >= nw) continue;
      int b = fb - nw;
      if (!is_awake(b)) continue;
      ++eligible;
      bodies |= 1u << b;
      seen.set(p);
    }
    bool need = false;
    for (int b = 0; b < kMaxBodies; ++b) {
      if (b < sc.nb && ((bodies >> b) & 1u)) {
        uint32_t m = toi_prefilter(b);
        for (int k = 0; k < ncl; ++k) {
          int p = clist[k];
          if (seen.test(p) && sc.pair[p].fb - nw == b && ((m >> sc.pair[p].fa) & 1u)) need = true;
        }
      }
    }
    if (!need) cnt[BLCD_CNT_TOI_CALLS] += (uint32_t)eligible;
    return need;
  }

  BLCD_HD void draw_action(float* action) {
    for (int k = 0; k < sc.A; ++k) action[k] = (float)rng.uniform(-1.0, 1.0);
  }

  // ---- reset (world_env.py:197-385) ----------------------------------------------------------------------------------
  BLCD_HD static double mapto(double x, double lo, double hi) { return ((x + 1.0) / 2.0 * (hi - lo)) + lo; }
  BLCD_HD static double rmapto(double x, double lo, double hi) { return ((x - lo) / (hi - lo) * 2.0) + -1.0; }

  // fresh b2World with every dynamic body created at pose[b] = (x, y, angle)
  BLCD_HD void build_fresh(const float (*pose)[3]) {
    awake = (1u << sc.nb) - 1u;
    newFixture = true;
    inv_dt0 = 0.0f;
    ep_t = 0;
    ncl = 0;
    slotUsed = 0u;
    moved = 0u;
    for (int k = 0; k < kMaxPairs; ++k) pslot[k] = -1;
    for (int s = 0; s < sc.maxm; ++s) g.u(slot_base(s) + S_HDR) = kSlotFree;
    for (int k = 0; k < BLCD_N_COUNTERS; ++k) cnt[k] = 0u;
    for (int b = 0; b < kMaxBodies; ++b) {
      if (b < sc.nb) {
        Xf t;
        t.p = mk(pose[b][0], pose[b][1]);
        t.q = rot_of(pose[b][2]);
        xf[b] = t;
        c[b] = xmul(t, lc_of(b)); c0[b] = c[b];
        a[b] = pose[b][2]; a0[b] = a[b];
        v[b] = mk(0.0f, 0.0f); w[b] = 0.0f; sleepT[b] = 0.0f; alpha0[b] = 0.0f;
        Box bb = shape_aabb(bshape(b), t);
        V2 r = mk(kAabbExtension, kAabbExtension);
        fat[b].lo = bb.lo - r;
        fat[b].hi = bb.hi + r;
      }
    }
    for (int j = 0; j < sc.nj; ++j) {
      int h = kHotJoint * j;
      jr[h + J_IX] = 0.0f; jr[h + J_IY] = 0.0f; jr[h + J_IZ] = 0.0f; jr[h + J_MI] = 0.0f; jr[h + J_MS] = 0.0f;
      jru(h + J_PK) = 0u;
      // pybox2d's revoluteJointDef(bodyA=, bodyB=, ...) sets referenceAngle = bodyB.angle - bodyA.angle when the joint is
      // defined (world_env.py:255-266): limits are relative to the pose the robot is assembled in (oracle/b2_oracle.cpp)
      jr[h + J_REF] = pose[sc.joint[j].b][2] - pose[sc.joint[j].a][2];
    }
  }

  // b2Body::SetTransform
  BLCD_HD void set_transform(int b, V2 position, float angle) {
    Xf t;
    t.q = rot_of(angle);
    t.p = position;
    xf[b] = t;
    c[b] = xmul(t, lc_of(b)); a[b] = angle;
    c0[b] = c[b]; a0[b] = angle;
    sync_fixture(b, t);
  }

  BLCD_HDN void reset(const float* full_state) {
    float pose[kMaxBodies][3];
    double angle64[kMaxBodies];
    const double W = sc.world_w, H = sc.world_h;
    variant = 0u;
    for (int b = 0; b < sc.nb; ++b) {
      const DBody& bd = sc.body[b];
      if (bd.role == BLCD_ROLE_ROOT) {
        double rangex = 1.0 - (2.0 * bd.extent / W), rangey = 1.0 - (2.0 * bd.extent / H);
        double x = mapto(rng.uniform(-rangex, rangex), 0.0, W);
        double y = mapto(rng.uniform(-rangey, -rangey), 0.0, H);
        double s = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
        double cc = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
        double ang = atan2(s, cc);
        if (!bd.rand_angle) ang = 0.0;
        angle64[b] = ang;
        pose[b][0] = (float)x; pose[b][1] = (float)y; pose[b][2] = (float)ang;
      } else if (bd.role == BLCD_ROLE_CHILD) {
        double mangle = angle64[bd.root] + bd.joint_angle;
        mangle = atan2(sin(mangle), cos(mangle));
        angle64[b] = mangle;
        double pangle = angle64[bd.parent];
        double aax = cos(pangle) * bd.anchor_a[0] - sin(pangle) * bd.anchor_a[1];
        double aay = sin(pangle) * bd.anchor_a[0] + cos(pangle) * bd.anchor_a[1];
        double abx = cos(mangle) * bd.anchor_b[0] - sin(mangle) * bd.anchor_b[1];
        double aby = sin(mangle) * bd.anchor_b[0] + cos(mangle) * bd.anchor_b[1];
        float px = pose[bd.parent][0] + (float)aax, py = pose[bd.parent][1] + (float)aay;
        px = px - (float)abx; py = py - (float)aby;
        pose[b][0] = px; pose[b][1] = py; pose[b][2] = (float)mangle;
      } else {
        if (bd.nvar > 1) variant |= (rng.next() & 1u) << b;
        double rangex = 1.0 - (2.0 * bd.extent / W), rangey = 1.0 - (2.0 * bd.extent / H);
        double x = mapto(rng.uniform(-rangex, rangex), 0.0, W);
        double y = sc.has_robot ? mapto(rng.uniform(-rangey, -0.25), 0.0, H) : mapto(rng.uniform(-rangey, rangey), 0.0, H);
        double ang = 0.0;
        if (bd.rand_angle) {
          double s = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
          double cc = mapto(rng.uniform(-1.0, 1.0), -1.0, 1.0);
          ang = atan2(s, cc);
        }
        pose[b][0] = (float)x; pose[b][1] = (float)y; pose[b][2] = (float)ang;
      }
    }
    build_fresh(pose);
    if (full_state) {
      for (int pass = 0; pass < 2; ++pass) {
        for (int b = 0; b < sc.nb; ++b) {
          const DBody& bd = sc.body[b];
          if ((bd.role == BLCD_ROLE_OBJECT) != (pass == 0)) continue;
          double x = mapto((double)full_state[bd.obs[0]], 0.0, W);
          double y = mapto((double)full_state[bd.obs[1]], 0.0, H);
          double ang = atan2((double)full_state[bd.obs[3]], (double)full_state[bd.obs[2]]);
          set_transform(b, mk((float)x, (float)y), a[b]);
          set_transform(b, xf[b].p, (float)ang);
        }
      }
      moved = 0u;
    }
  }

  // ---- observation (world_env.py:387-429, 460-512) -------------------------------------------------------------------
  BLCD_HD void obs_body(int b, float out[4]) const {
    out[0] = (float)rmapto((double)xf[b].p.x, 0.0, (double)sc.world_w);
    out[1] = (float)rmapto((double)xf[b].p.y, 0.0, (double)sc.world_h);
    out[2] = (float)cos((double)a[b]);
    out[3] = (float)sin((double)a[b]);
  }

  BLCD_HD RowMask lcd_row(int R) const {  // output row R (0 = top of the world)
    int y = sc.lcd_h - 1 - R;
    RowMask ink = 0u;
    for (int b = 0; b < sc.nb; ++b)
      ink |= body_row(bshape(b), xf[b].p.x, xf[b].p.y, xf[b].q.s, xf[b].q.c, y, sc.world_w, sc.lcd_w, sc.lcd_h, sc.rules);
    return row_bits_from_ink(ink, sc.lcd_w);
  }
};
#undef sc

}  // namespace BLCD_NS
