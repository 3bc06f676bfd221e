// boxlcd_b200.cu -- CUDA kernels (sm_100a) and the C ABI of libboxlcd_b200.so (include/boxlcd_b200.h).
//
// Kernels (one thread per world unless noted; shared memory laid out [word][thread], see blcd_world.cuh):
//   k_init          mark all manifold slots free after the state buffer was zeroed
//   k_reset         WorldEnv.reset          (world_env.py:306-385)
//   k_set_bodies    fresh world at given poses / velocities (single-step parity protocol)
//   k_step          WorldEnv.step           (world_env.py:431-458), optionally followed by _get_obs
//   k_rollout       examples/collect.py:31-39 inner loop, T steps per launch, actions from the per-world Philox stream
//   k_observe       WorldEnv._get_obs       (world_env.py:387-429)
//   k_render_poses  WorldEnv.lcd_render     (world_env.py:460-512), one thread per (world, frame row)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
#include "blcd_pipeline.cuh"

// This file is compiled once per scene-size profile (blcd_profile.h).  Its C entry points carry the profile's prefix
// (blcd_small_* / blcd_large_*, declared in blcd_profile_api.h); blcd_dispatch.cpp owns the public blcd_* symbols.
#define BLCD_CAT2(a, b) a##b
#define BLCD_CAT(a, b) BLCD_CAT2(a, b)
#if defined(BLCD_PROFILE_LARGE)
#define BLCD_P(name) BLCD_CAT(blcd_large_, name)
#define BLCD_PENV blcd_large_env
#else
#define BLCD_P(name) BLCD_CAT(blcd_small_, name)
#define BLCD_PENV blcd_small_env
#endif
#include "blcd_profile_api.h"

using namespace BLCD_NS;

namespace {

int fail(const std::string& msg) { return blcd_fail_msg(msg.c_str()); }
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));              \
  } while (0)

struct OutPtrs {
  float* full_state;   // [N, S] or [N, T, S]
  float* proprio;      // [N, P]
  uint32_t* lcd_bits;  // [N, H] or [N, T, H]
  uint8_t* lcd_bool;   // [N, H, W]
  uint8_t* done;       // [N]
  float* actions;      // [N, A] or [N, T, A]
};

template <int BLOCK>
__device__ __forceinline__ const DScene& stage_scene(const DScene* scene_g, unsigned char* smem_raw) {
  // the scene table is read by every thread with divergent indices: keep one copy per block in shared memory
  DScene* sc = reinterpret_cast<DScene*>(smem_raw);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(scene_g);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sc);
  for (int i = threadIdx.x; i < (int)(sizeof(DScene) / 4); i += BLOCK) dst[i] = src[i];
  __syncthreads();
  return *sc;
}


template <int BLOCK>
__device__ __forceinline__ float* hot_base(unsigned char* smem_raw) {
  return reinterpret_cast<float*>(smem_raw + kSceneBytes) + threadIdx.x;
}

template <int BLOCK>
__device__ __forceinline__ void write_obs(const Sim<BLOCK>& sim, const DScene& sc, const OutPtrs& o, int64_t row) {
  // row = world index (or world * T + t for rollouts)
  if (o.full_state || o.proprio) {
    float fs[kMaxObs];
    for (int b = 0; b < kMaxBodies; ++b) {
      if (b < sc.nb) {
        float ob[4];
        sim.obs_body(b, ob);
        for (int k = 0; k < 4; ++k) fs[sc.body[b].obs[k]] = ob[k];
      }
    }
    // dataset rows are written once and never read back by the kernel: streaming stores (evict-first) keep them from
    // pushing the thread-local solver records out of L2
    if (o.full_state) {
      float4* dst = reinterpret_cast<float4*>(o.full_state + row * sc.S);   // S = 4 * n_bodies, rows are 16-byte aligned
      for (int i = 0; i < sc.S / 4; ++i) __stcs(dst + i, make_float4(fs[4 * i], fs[4 * i + 1], fs[4 * i + 2], fs[4 * i + 3]));
    }
    if (o.proprio) {
      if (sc.P == 0) o.proprio[row] = 0.0f;
      for (int i = 0; i < sc.P; ++i) __stcs(o.proprio + row * sc.P + i, fs[sc.pobs[i]]);
    }
  }
  if (o.lcd_bits || o.lcd_bool) {
    BodyPx bp[kMaxBodies];   // vertex transform + fp64 metre->pixel scaling once per body, not once per row
    for (int b = 0; b < kMaxBodies; ++b)
      if (b < sc.nb) body_px(bp[b], sim.bshape(b), sim.xf[b].p.x, sim.xf[b].p.y, sim.xf[b].q.s, sim.xf[b].q.c, sc.world_w, sc.lcd_w);
    const int lw = row_words(sc.lcd_w);
    for (int R = 0; R < sc.lcd_h; ++R) {
      const int y = sc.lcd_h - 1 - R;
      RowMask ink = 0u;
      for (int b = 0; b < kMaxBodies; ++b)
        if (b < sc.nb) ink |= body_px_row(bp[b], y, sc.lcd_w, sc.lcd_h, sc.rules);
      RowMask bits = row_bits_from_ink(ink, sc.lcd_w);
      if (o.lcd_bits) {
        if (sizeof(RowMask) == 4) __stcs(o.lcd_bits + row * sc.lcd_h + R, (uint32_t)bits);
        else for (int k = 0; k < lw; ++k) __stcs(o.lcd_bits + (row * sc.lcd_h + R) * lw + k, row_word(bits, k));
      }
      if (o.lcd_bool)
        for (int x = 0; x < sc.lcd_w; ++x) o.lcd_bool[(row * sc.lcd_h + R) * sc.lcd_w + x] = (uint8_t)((bits >> x) & 1u);
    }
  }
  if (o.done) o.done[row] = (uint8_t)(sim.ep_t >= sc.ep_len);
}

__global__ void k_init(const DScene* scene_g, uint32_t* state, int64_t n) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  for (int s = 0; s < scene_g->maxm; ++s) state[(int64_t)(scene_g->off_slots + kSlotWords * s) * n + w] = kSlotFree;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_reset(const DScene* scene_g, uint32_t* state, int64_t n, uint64_t seed, int64_t world_offset,
                                                  const int64_t* idx, int64_t n_idx, const float* full_state) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<BLOCK>(scene_g, smem_raw);
  int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n_idx) return;
  int64_t w = idx ? idx[i] : i;
  if (w < 0 || w >= n) return;
  Sim<BLOCK> sim(sc, hot_base<BLOCK>(smem_raw), state, n, w);
  sim.load(seed, world_offset + w);
  float fs[kMaxObs];
  if (full_state)
    for (int k = 0; k < sc.S; ++k) fs[k] = full_state[i * sc.S + k];
  sim.reset(full_state ? fs : nullptr);
  sim.store();
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_set_bodies(const DScene* scene_g, uint32_t* state, int64_t n, uint64_t seed, int64_t world_offset,
                                                       const float* bodies, const uint32_t* variants) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<BLOCK>(scene_g, smem_raw);
  int64_t w = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (w >= n) return;
  Sim<BLOCK> sim(sc, hot_base<BLOCK>(smem_raw), state, n, w);
  sim.load(seed, world_offset + w);
  sim.variant = variants ? (variants[w] & kBodyMask) : 0u;
  float pose[kMaxBodies][3];
  const float* src = bodies + w * sc.nb * BLCD_BODY_STATE;
  for (int b = 0; b < sc.nb; ++b) { pose[b][0] = src[b * 6]; pose[b][1] = src[b * 6 + 1]; pose[b][2] = src[b * 6 + 2]; }
  sim.build_fresh(pose);
  for (int b = 0; b < sc.nb; ++b) { sim.v[b] = mk(src[b * 6 + 3], src[b * 6 + 4]); sim.w[b] = src[b * 6 + 5]; }
  sim.store();
}

__global__ void k_get_bodies(const DScene* scene_g, const uint32_t* state, int64_t n, float* bodies) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const DScene& sc = *scene_g;
  const float* sf = reinterpret_cast<const float*>(state);
  for (int b = 0; b < sc.nb; ++b) {
    int o = kBodyWords * b;
    float a = sf[(int64_t)(o + 2) * n + w];
    float* dst = bodies + (w * sc.nb + b) * BLCD_BODY_STATE;
    dst[0] = sf[(int64_t)(o + 11) * n + w]; dst[1] = sf[(int64_t)(o + 12) * n + w]; dst[2] = a;
    dst[3] = sf[(int64_t)(o + 3) * n + w]; dst[4] = sf[(int64_t)(o + 4) * n + w]; dst[5] = sf[(int64_t)(o + 5) * n + w];
  }
}

// failure detection: worlds whose body state stopped being finite (a diverged solve); flags[w] = 1 for those
__global__ void k_check_finite(const DScene* scene_g, const uint32_t* state, int64_t n, uint8_t* flags, unsigned long long* count) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const DScene& sc = *scene_g;
  const float* sf = reinterpret_cast<const float*>(state);
  bool ok = true;
  for (int b = 0; b < sc.nb; ++b)
    for (int k = 0; k < 6; ++k) ok = ok && isfinite(sf[(int64_t)(kBodyWords * b + k) * n + w]);
  if (flags) flags[w] = ok ? 0 : 1;
  if (!ok) atomicAdd(count, 1ull);
}

// b2Transform of every dynamic body exactly as the simulation holds it (position, sincosf(angle)): what lcd_render consumes
__global__ void k_get_poses(const DScene* scene_g, const uint32_t* state, int64_t n, float* poses, uint32_t* variants) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const DScene& sc = *scene_g;
  const float* sf = reinterpret_cast<const float*>(state);
  for (int b = 0; b < sc.nb; ++b) {
    int o = kBodyWords * b;
    Rot q = rot_of(sf[(int64_t)(o + 2) * n + w]);
    float* dst = poses + (w * sc.nb + b) * 4;
    dst[0] = sf[(int64_t)(o + 11) * n + w]; dst[1] = sf[(int64_t)(o + 12) * n + w]; dst[2] = q.s; dst[3] = q.c;
  }
  if (variants) variants[w] = kVariantInFlags ? ((state[(int64_t)sc.off_misc * n + w] >> kVariantShift) & kBodyMask) : state[(int64_t)(sc.off_misc + 4) * n + w];
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK >= 256 ? 1 : 256 / BLOCK) k_step(const DScene* scene_g, uint32_t* state, int64_t n, uint64_t seed, int64_t world_offset,
                                                 const float* actions, int n_steps, OutPtrs out, int64_t w_begin, int64_t w_end) {
  // worlds [w_begin, w_end) of the handle's n (blcd_step_host pipelines sub-ranges over several streams)
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<BLOCK>(scene_g, smem_raw);
  int64_t w = w_begin + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (w >= w_end) return;
  Sim<BLOCK> sim(sc, hot_base<BLOCK>(smem_raw), state, n, w, w_end);
  sim.load(seed, world_offset + w);
  float act[kMaxObs];
  for (int t = 0; t < n_steps; ++t) {
    if (actions) { for (int k = 0; k < sc.A; ++k) act[k] = actions[w * sc.A + k]; }
    else sim.draw_action(act);
    if (out.actions)
      for (int k = 0; k < sc.A; ++k) out.actions[w * sc.A + k] = act[k];
    sim.env_step(act);
  }
  OutPtrs o = out;
  o.actions = nullptr;
  write_obs<BLOCK>(sim, sc, o, w);
  sim.store();
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, BLOCK >= 256 ? 1 : 256 / BLOCK) k_rollout(const DScene* scene_g, uint32_t* state, int64_t n, uint64_t seed, int64_t world_offset,
                                                    int T, OutPtrs out) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<BLOCK>(scene_g, smem_raw);
  int64_t w = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (w >= n) return;
  Sim<BLOCK> sim(sc, hot_base<BLOCK>(smem_raw), state, n, w);
  sim.load(seed, world_offset + w);
  float act[kMaxObs];
  sim.ph_start();
  for (int t = 0; t < T; ++t) {
    int64_t row = w * T + t;
    sim.ph(4);
    write_obs<BLOCK>(sim, sc, out, row);   // obs_t is recorded before action t (collect.py:33-39)
    sim.ph(6);
    sim.draw_action(act);
    if (out.actions)
      for (int k = 0; k < sc.A; ++k) __stcs(out.actions + row * sc.A + k, act[k]);
    sim.env_step(act);
  }
  sim.store();
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_observe(const DScene* scene_g, uint32_t* state, int64_t n, OutPtrs out, int64_t w_begin, int64_t w_end) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<BLOCK>(scene_g, smem_raw);
  int64_t w = w_begin + (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (w >= w_end) return;
  Sim<BLOCK> sim(sc, hot_base<BLOCK>(smem_raw), state, n, w);
  sim.load(0, 0);
  write_obs<BLOCK>(sim, sc, out, w);
}

// ---- phase pipeline kernels (blcd_pipeline.cuh): one sub-step = pre -> vel -> pos -> post -> toi over worlds [w_begin, w_end) --
// All of them: one thread per world, kPipeBlock threads per block (body rows in shared memory [word][thread]), no
// block-level synchronisation after the scene table is staged, so blocks and warps drift freely.
constexpr int kPipeBlock = 128;
// resident blocks per SM each phase kernel is compiled for (register cap = 65536 / (128 x blocks)) and given shared memory for
constexpr int kPreBlocks = 4, kVelBlocks = 4, kPosBlocks = 4, kPostBlocks = 4;
// The velocity kernel takes its worlds in SORTED order: k_pipe_pre files every world under the key (touching contacts, how
// many of them have two points), k_pipe_vel walks the bins from the busiest key down.  A warp then holds 32 worlds with the
// same contact count, so the contact part of a velocity sweep runs with full warps instead of the warp's busiest lane
// setting the trip count for everyone (7 of 32 lanes active before).  Any lane can take any world because everything the
// velocity loop touches comes from the scratch area.
constexpr int kVelBinsC = 8, kVelBinsP = 4, kVelBins = kVelBinsC * kVelBinsP;
// The position kernel hands out worlds longest-first: sweep counts are heavy-tailed (mean ~8, up to 60) and a world tends to
// need about as many sweeps as in its previous sub-step, so worlds are also filed by that count (kPosBins bins of 4 sweeps) and
// fetched from the top bin down -- the stragglers start early and the cheap worlds fill the tail of the launch.
constexpr int kPosBins = 16, kAllBins = kVelBins + kPosBins;
// atomicAdd(counter + key, 1) for every calling lane, aggregated per warp: lanes with the same key elect a leader that adds the
// group size once; each lane gets its own slot.  (One atomic per world on a handful of hot addresses serialises at L2.)
__device__ __forceinline__ uint32_t warp_agg_inc(uint32_t* counters, int key) {
  const unsigned peers = __match_any_sync(__activemask(), key);
  const int leader = __ffs(peers) - 1, lane = (int)(threadIdx.x & 31u);
  uint32_t base = 0u;
  if (lane == leader) base = atomicAdd(counters + key, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
}

__device__ __forceinline__ int vel_bin_key(int nc, int n2) {
  return (nc < kVelBinsC ? nc : kVelBinsC - 1) * kVelBinsP + (n2 < kVelBinsP ? n2 : kVelBinsP - 1);
}

// mode 0: blcd_step (actions given or drawn; out.actions [N, A]); mode 1: blcd_rollout (obs_t and a_t recorded at row w T + t)
__global__ void __launch_bounds__(kPipeBlock, kPreBlocks) k_pipe_pre(const DScene* scene_g, uint32_t* state, uint32_t* scratch, int64_t n, uint64_t seed, int64_t world_offset,
                                                            const float* actions, int mode, int T, int t, int first, OutPtrs out,
                                                            int64_t w_begin, int64_t w_end, uint32_t* toi_count, unsigned long long* pos_next,
                                                            uint32_t* bin_count, uint32_t* bin_list, int file_prev) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<kPipeBlock>(scene_g, smem_raw);
  if (blockIdx.x == 0 && threadIdx.x == 0) { *toi_count = 0u; *pos_next = 0ull; }   // this sub-step's TOI list and position work counter start empty
  int64_t w = w_begin + (int64_t)blockIdx.x * kPipeBlock + threadIdx.x;
  if (w >= w_end) return;
  Sim<kPipeBlock> sim(sc, hot_base<kPipeBlock>(smem_raw), state, n, w, w_end);
  sim.attach_scratch(scratch, w);
  sim.load(seed, world_offset + w);
  float act[kMaxObs];
  if (first) {
    if (mode == 1) {
      int64_t row = w * T + t;
      write_obs<kPipeBlock>(sim, sc, out, row);   // obs_t is recorded before action t (collect.py:33-39)
      sim.draw_action(act);
      if (out.actions)
        for (int k = 0; k < sc.A; ++k) __stcs(out.actions + row * sc.A + k, act[k]);
    } else {
      if (actions) { for (int k = 0; k < sc.A; ++k) act[k] = actions[w * sc.A + k]; }
      else sim.draw_action(act);
      if (out.actions)
        for (int k = 0; k < sc.A; ++k) out.actions[w * sc.A + k] = act[k];
    }
  }
  pipe_pre(sim, first != 0, act);
  int n2 = 0;
  for (int k = 0; k < sim.nc; ++k) n2 += ((sim.cru(kHotCon * k + C_PK) >> 10) & 3u) == 2u;
  const int key = vel_bin_key(sim.nc, n2);
  bin_list[(int64_t)key * n + w_begin + warp_agg_inc(bin_count, key)] = (uint32_t)(w - w_begin);
  if (file_prev) {   // the longest-first order of the position kernel (BLCD_POS_SORT=1): by the sweeps needed last time
    const uint32_t prev = sim.x.u(pipe_prev_sweeps_word(sc));
    const int pkey = kVelBins + (int)(prev / 4u < (uint32_t)kPosBins ? prev / 4u : (uint32_t)kPosBins - 1u);
    bin_list[(int64_t)pkey * n + w_begin + warp_agg_inc(bin_count, pkey)] = (uint32_t)(w - w_begin);
  }
}

__global__ void __launch_bounds__(kPipeBlock, kVelBlocks) k_pipe_vel(const DScene* scene_g, uint32_t* state, uint32_t* scratch, int64_t n, int64_t w_begin, int64_t w_end,
                                                                     const uint32_t* bin_count, const uint32_t* bin_list) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<kPipeBlock>(scene_g, smem_raw);
  int64_t i = (int64_t)blockIdx.x * kPipeBlock + threadIdx.x;
  if (i >= w_end - w_begin) return;
  // position i of the sorted order: bins from the busiest key down (the longest-running warps start first)
  int64_t w = w_begin + i;
  if (bin_list) {
    int key = kVelBins - 1;
    for (; key > 0; --key) {
      const int64_t c = (int64_t)bin_count[key];
      if (i < c) break;
      i -= c;
    }
    w = w_begin + (int64_t)bin_list[(int64_t)key * n + w_begin + i];
  }
  Sim<kPipeBlock> sim(sc, hot_base<kPipeBlock>(smem_raw), state, n, w, w_end);
  sim.attach_scratch(scratch, w);
  pipe_vel(sim);
}

// Position iterations with LANES DECOUPLED FROM WORLDS.  A world needs anything from 0 to 60 sweeps (mean ~8-14): with one
// world per lane a warp runs until its slowest world is done (4 of 32 lanes active on average).  Here the kernel is
// persistent and every loop trip runs ONE sweep for whatever world each lane currently holds; a lane whose world has
// converged writes it back and fetches the next unprocessed world from a global counter.  Everything a sweep needs is in
// the lane's own shared-memory column / local records, filled from the scratch area, so any lane can take any world.
__global__ void __launch_bounds__(kPipeBlock, kPosBlocks) k_pipe_pos(const DScene* scene_g, uint32_t* state, uint32_t* scratch, int64_t n, int64_t w_begin, int64_t w_end,
                                                                     unsigned long long* next, int refill, const uint32_t* bin_count, const uint32_t* bin_list, int sort_mode) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<kPipeBlock>(scene_g, smem_raw);
  Sim<kPipeBlock> sim(sc, hot_base<kPipeBlock>(smem_raw), state, n, w_begin);
  const long long count = (long long)(w_end - w_begin);
  bool have = false, exhausted = false;
  int it = 0;
  for (;;) {
    // Refill in batches: fetching a world (filling its rows and records from the scratch area) is a few hundred instructions
    // that only the fetching lanes execute, so lanes that ran out wait until `refill` of them are idle (or nobody has work)
    // and then fetch together.
    const unsigned idle = __ballot_sync(0xFFFFFFFFu, !have && !exhausted);
    if (idle && (__popc(idle) >= refill || !__any_sync(0xFFFFFFFFu, have))) {
      while (!have && !exhausted) {
        // one atomic per warp and round: the fetching lanes take consecutive indices
        const unsigned takers = __activemask();
        const int lane = (int)(threadIdx.x & 31u), leader = __ffs(takers) - 1;
        unsigned long long base = 0ull;
        if (lane == leader) base = atomicAdd(next, (unsigned long long)__popc(takers));
        base = __shfl_sync(takers, base, leader);
        long long idx = (long long)base + __popc(takers & ((1u << lane) - 1u));
        if (idx >= count) { exhausted = true; break; }
        int64_t w = w_begin + idx;
        if (bin_list && sort_mode == 1) {   // position idx of the longest-first order (sweeps needed last time)
          int key = kPosBins - 1;
          for (; key > 0; --key) {
            const long long c = (long long)bin_count[kVelBins + key];
            if (idx < c) break;
            idx -= c;
          }
          w = w_begin + (int64_t)bin_list[(int64_t)(kVelBins + key) * n + w_begin + idx];
        } else if (bin_list) {              // the velocity kernel's order: worlds with the same number of contacts arrive together
          int key = kVelBins - 1;
          for (; key > 0; --key) {
            const long long c = (long long)bin_count[key];
            if (idx < c) break;
            idx -= c;
          }
          w = w_begin + (int64_t)bin_list[(int64_t)key * n + w_begin + idx];
        }
        sim.g.p = state + w;
        sim.attach_scratch(scratch, w);
        pipe_pos_begin(sim);
        it = 0;
        have = sim.nIslands > 0 && sc.pos_iters > 0;
        if (!have) pipe_pos_end(sim);
      }
    }
    if (!__any_sync(0xFFFFFFFFu, have)) {
      if (__all_sync(0xFFFFFFFFu, exhausted)) break;   // the whole warp is out of work and the queue is empty
      continue;
    }
    if (have) {
      bool done = sim.solve_position_sweep<true>();
      ++it;
      if (done || it >= sc.pos_iters) { pipe_pos_end(sim, it); have = false; }
    }
  }
}

__global__ void __launch_bounds__(kPipeBlock, kPostBlocks) k_pipe_post(const DScene* scene_g, uint32_t* state, uint32_t* scratch, int64_t n, uint64_t seed, int64_t world_offset,
                                                             int64_t w_begin, int64_t w_end, uint32_t* toi_count, uint32_t* toi_list, uint32_t* bin_count) {
  unsigned char* smem_raw = blcd_smem;
  const DScene& sc = stage_scene<kPipeBlock>(scene_g, smem_raw);
  if (blockIdx.x == 0 && threadIdx.x < kAllBins) bin_count[threadIdx.x] = 0u;   // consumed by k_pipe_vel; refilled by the next k_pipe_pre
  int64_t w = w_begin + (int64_t)blockIdx.x * kPipeBlock + threadIdx.x;
  if (w >= w_end) return;
  Sim<kPipeBlock> sim(sc, hot_base<kPipeBlock>(smem_raw), state, n, w, w_end);
  sim.attach_scratch(scratch, w);
  sim.load(seed, world_offset + w);
  if (pipe_post(sim)) toi_list[warp_agg_inc(toi_count, 0)] = (uint32_t)(w - w_begin);
}

// One warp per block: a world's SolveTOI takes anything from one pre-filtered scan to eight events, and a block holds its
// registers and shared memory until its slowest world is done; with one-warp blocks the resources of finished warps return
// to the SM at once (LuxoCube +5 %, Urchin unchanged).
constexpr int kToiBlock = 32;
__global__ void __launch_bounds__(kToiBlock, 16) k_pipe_toi(const DScene* scene_g, uint32_t* state, uint32_t* scratch, int64_t n, uint64_t seed, int64_t world_offset,
                                                            int64_t w_begin, const uint32_t* toi_count, const uint32_t* toi_list) {
  unsigned char* smem_raw = blcd_smem;
  int64_t i = (int64_t)blockIdx.x * kToiBlock + threadIdx.x;
  if ((int64_t)blockIdx.x * kToiBlock >= (int64_t)*toi_count) return;   // nothing for this block: leave before staging the scene
  const DScene& sc = stage_scene<kToiBlock>(scene_g, smem_raw);
  if (i >= (int64_t)*toi_count) return;
  int64_t w = w_begin + (int64_t)toi_list[i];
  Sim<kToiBlock> sim(sc, hot_base<kToiBlock>(smem_raw), state, n, w);
  sim.attach_scratch(scratch, w);
  sim.load(seed, world_offset + w);
  pipe_toi(sim);
}

// lcd_render from poses.  One thread per (frame, row): lanes 0..H-1 of consecutive frames write consecutive words, so the
// stores are fully coalesced and need no atomics.  The H lanes of a frame first share the per-frame set-up through shared
// memory -- each vertex is transformed (fp32, no FMA) and scaled to pixels (fp64, truncation) ONCE per frame by one lane
// instead of once per row -- then every lane scans its own row over all bodies.
constexpr int kRenderThreads = 256;
typedef std::conditional<sizeof(RowMask) == 4, unsigned int, unsigned long long>::type RowInk;   // atomicOr operand type
constexpr int kRenderMaxFrames = 32;   // frames per block (bounds the BodyPx staging area in shared memory)
__host__ __device__ inline int render_frames_per_block(int lcd_h) { int f = kRenderThreads / lcd_h; return f < kRenderMaxFrames ? f : kRenderMaxFrames; }
// One launch renders the column window [x_off, x_off + win_w) of every frame (win_w <= bits of a RowMask) into words
// word_off .. of each output row of row_stride words; frames wider than a RowMask take one launch per window.
__global__ void __launch_bounds__(kRenderThreads) k_render_poses(const DScene* scene_g, const float* poses, const uint32_t* variants, int64_t n,
                                                                  int lcd_w, int lcd_h, uint32_t* bits, int x_off, int win_w, int row_stride, int word_off) {
  extern __shared__ __align__(16) unsigned char rsm[];
  DScene* scp = reinterpret_cast<DScene*>(rsm);
  BodyPx* bp_all = reinterpret_cast<BodyPx*>(rsm + kSceneBytes);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(scene_g);
    uint32_t* dst = reinterpret_cast<uint32_t*>(scp);
    for (int i = threadIdx.x; i < (int)(sizeof(DScene) / 4); i += kRenderThreads) dst[i] = src[i];
    __syncthreads();
  }
  const DScene& sc = *scp;
  const int fpb = render_frames_per_block(lcd_h);
  const int f = threadIdx.x / lcd_h, R = threadIdx.x - f * lcd_h;
  const int64_t w = (int64_t)blockIdx.x * fpb + f;
  const bool live = f < fpb && w < n;
  BodyPx* bp = bp_all + f * kMaxBodies;
  const uint32_t variant = (live && variants) ? variants[w] : 0u;
  if (live) {
    const double ww = (double)sc.world_w, lw = (double)lcd_w;
    for (int item = R; item < sc.nb * BLCD_MAX_VERTS; item += lcd_h) {
      const int b = item / BLCD_MAX_VERTS, vi = item - b * BLCD_MAX_VERTS;
      const DShape& sh = sc.body[b].shape[(variant >> b) & 1u];
      const float* p = poses + (w * sc.nb + b) * 4;
      if (sh.type == SH_CIRCLE) {
        if (vi == 0) {
          const double r = (double)sh.radius;
          bp[b].kind = SH_CIRCLE;
          bp[b].x0 = to_px((double)p[0] - r, ww, lw); bp[b].y0 = to_px((double)p[1] - r, ww, lw);
          bp[b].x1 = to_px((double)p[0] + r, ww, lw); bp[b].y1 = to_px((double)p[1] + r, ww, lw);
          bp[b].P.n = 0;
        }
      } else if (vi < sh.count) {
        const float px = p[0], py = p[1], sn = p[2], cs = p[3], vx = sh.v[vi].x, vy = sh.v[vi].y;
        float wx = BLCD_FADD(BLCD_FSUB(BLCD_FMUL(cs, vx), BLCD_FMUL(sn, vy)), px);
        float wy = BLCD_FADD(BLCD_FADD(BLCD_FMUL(sn, vx), BLCD_FMUL(cs, vy)), py);
        bp[b].P.x[vi] = to_px((double)wx, ww, lw);
        bp[b].P.y[vi] = to_px((double)wy, ww, lw);
        if (vi == 0) { bp[b].kind = SH_POLY; bp[b].P.n = sh.count; }
      }
    }
  }
  __syncthreads();
  if (live) {
    for (int b = R; b < sc.nb; b += lcd_h) {               // row bounds of each polygon, one lane per body
      if (bp[b].kind != SH_CIRCLE) {
        int ylo = bp[b].P.y[0], yhi = bp[b].P.y[0];
        for (int i = 1; i < bp[b].P.n; ++i) { ylo = min(ylo, bp[b].P.y[i]); yhi = max(yhi, bp[b].P.y[i]); }
        bp[b].y0 = ylo; bp[b].y1 = yhi;
      }
    }
  }
  // Stage 2: work items are the (body, row) pairs that can contain ink.  They are dealt round-robin to the frame's lanes,
  // so every lane scans a row that matters instead of most lanes rejecting most bodies; spans are OR-ed into the
  // frame's rows in shared memory.
  RowInk* rowink = reinterpret_cast<RowInk*>(bp_all + fpb * kMaxBodies) + f * lcd_h;
  if (f < fpb) rowink[R] = 0u;
  __syncthreads();
  if (live) {
    int total = 0;
    for (int b = 0; b < sc.nb; ++b) {
      int lo = max(bp[b].y0, 0), hi = min(bp[b].y1, lcd_h - 1);
      total += max(hi - lo + 1, 0);
    }
    for (int item = R; item < total; item += lcd_h) {
      int rem = item, b = 0, lo = 0;
      for (;; ++b) {
        lo = max(bp[b].y0, 0);
        int cnt = max(min(bp[b].y1, lcd_h - 1) - lo + 1, 0);
        if (rem < cnt) break;
        rem -= cnt;
      }
      const int y = lo + rem;
      RowMask m = body_px_row(bp[b], y, win_w, lcd_h, sc.rules, x_off);
      if (m) atomicOr(&rowink[y], (RowInk)m);
    }
  }
  __syncthreads();
  if (!live) return;
  const RowMask out = row_bits_from_ink((RowMask)rowink[lcd_h - 1 - R], win_w);
  for (int k = 0, lw = row_words(win_w); k < lw; ++k) bits[(w * lcd_h + R) * row_stride + word_off + k] = row_word(out, k);
}

// lcd_render, second formulation: ONE LANE PER (frame, body), lanes ordered body-major inside a block of kR2Frames frames.
// Consecutive lanes then hold the SAME body of consecutive frames -- the same shape kind, vertex count and size, hence the
// same code path and nearly the same number of rows -- so warps stay converged through the scanline rules, where the
// (frame, row)-per-lane kernel above mixes circles, polygons and empty rows in every warp (7.9 of 32 lanes active).  A lane
// sets its body up once (vertex transform, metre -> pixel scaling), walks the body's rows and ORs each row's ink into the
// frame's row masks in shared memory; the block then writes the kR2Frames x H packed rows with coalesced stores.
// Poses come either from a pose array [n, nb, 4] (blcd_render_poses) or straight from the simulation state (the rollout
// pipeline renders obs_t this way: frame w goes to output frame w * frame_mul + frame_add).
constexpr int kR2Frames = 64, kR2Threads = 256;
__global__ void __launch_bounds__(kR2Threads, 4) k_render_bodies(const DScene* scene_g, const float* poses, const uint32_t* variants, const uint32_t* state, int64_t n_state,
                                                              int64_t w_begin, int64_t n, int lcd_w, int lcd_h, uint32_t* bits, int x_off, int win_w, int row_stride, int word_off,
                                                              int64_t frame_mul, int64_t frame_add) {
  extern __shared__ __align__(16) unsigned char rsm[];
  DScene* scp = reinterpret_cast<DScene*>(rsm);
  RowInk* rowink = reinterpret_cast<RowInk*>(rsm + kSceneBytes);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(scene_g);
    uint32_t* dst = reinterpret_cast<uint32_t*>(scp);
    for (int i = threadIdx.x; i < (int)(sizeof(DScene) / 4); i += kR2Threads) dst[i] = src[i];
    for (int i = threadIdx.x; i < kR2Frames * lcd_h; i += kR2Threads) rowink[i] = 0u;
    __syncthreads();
  }
  const DScene& sc = *scp;
  const int64_t f0 = (int64_t)blockIdx.x * kR2Frames;
  const float* sf = reinterpret_cast<const float*>(state);
  for (int item = threadIdx.x; item < kR2Frames * sc.nb; item += kR2Threads) {
    const int b = item / kR2Frames, f = item - b * kR2Frames;
    const int64_t i = f0 + f;
    if (i >= n) continue;
    float px, py, sn, cs;
    uint32_t variant;
    if (state) {
      const int64_t w = w_begin + i;
      const int o = kBodyWords * b;
      px = sf[(int64_t)(o + 11) * n_state + w]; py = sf[(int64_t)(o + 12) * n_state + w];
      Rot q = rot_of(sf[(int64_t)(o + 2) * n_state + w]);
      sn = q.s; cs = q.c;
      variant = kVariantInFlags ? ((state[(int64_t)sc.off_misc * n_state + w] >> kVariantShift) & kBodyMask) : state[(int64_t)(sc.off_misc + 4) * n_state + w];
    } else {
      const float* p = poses + (i * sc.nb + b) * 4;
      px = p[0]; py = p[1]; sn = p[2]; cs = p[3];
      variant = variants ? variants[i] : 0u;
    }
    BodyPxFast bp;
    body_px_fast(bp, sc.body[b].shape[(variant >> b) & 1u], px, py, sn, cs, sc.world_w, lcd_w);
    const int ylo = max(bp.b.y0, 0), yhi = min(bp.b.y1, lcd_h - 1);
    for (int y = ylo; y <= yhi; ++y) {
      RowMask m = body_px_row_fast(bp, y, win_w, lcd_h, sc.rules, x_off);
      if (m) atomicOr(&rowink[f * lcd_h + y], (RowInk)m);
    }
  }
  __syncthreads();
  const int lw = row_words(win_w);
  for (int j = threadIdx.x; j < kR2Frames * lcd_h; j += kR2Threads) {
    const int f = j / lcd_h, R = j - f * lcd_h;
    const int64_t i = f0 + f;
    if (i >= n) continue;
    const RowMask out = row_bits_from_ink((RowMask)rowink[f * lcd_h + (lcd_h - 1 - R)], win_w);
    const int64_t frame = i * frame_mul + frame_add;
    for (int k = 0; k < lw; ++k) __stcs(bits + (frame * lcd_h + R) * row_stride + word_off + k, row_word(out, k));
  }
}

}  // namespace

constexpr int kHostStreams = 8;
constexpr int kHostDepth = 4;     // host steps that may be in flight (blcd_step_host_async)

struct BLCD_PENV {
  DScene scene;
  DScene* scene_dev = nullptr;
  uint32_t* state = nullptr;
  int64_t n = 0, world_offset = 0;
  uint64_t seed = 0;
  int device = 0, block = 64;
  int64_t launches = 0;
  bool timing = false, render_attr_set = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = -1.0f;
  // pinned staging for blcd_step_host
  float* h_act = nullptr; float* h_fs = nullptr; uint32_t* h_bits = nullptr; uint8_t* h_done = nullptr;
  float* d_act = nullptr; float* d_fs = nullptr; uint32_t* d_bits = nullptr; uint8_t* d_done = nullptr;
  std::vector<std::pair<const void*, size_t>> pinned_user;   // caller buffers page-locked by blcd_pin_host
  cudaStream_t hstream[kHostStreams] = {};              // blcd_step_host's pipeline streams
  cudaEvent_t hev[kHostDepth][kHostStreams] = {};       // completion of host step s (ring) on each stream
  uint64_t host_submitted = 0, host_completed = 0;
  int host_chunks = 0, host_chunks_env = -1;
  // phase pipeline (blcd_pipeline.cuh)
  int pipeline = 0;                 // 1: blcd_step / blcd_rollout run the phase pipeline instead of the fused kernel
  uint32_t* scratch = nullptr;      // [scratch_words][n]
  uint32_t* toi_count = nullptr;    // [kHostStreams] one counter per concurrently processed world range
  uint32_t* toi_list = nullptr;     // [n] range-relative world indices that need SolveTOI in the current sub-step
  unsigned long long* pos_next = nullptr;   // [kHostStreams] work counters of the persistent position kernel
  uint32_t* bin_count = nullptr;    // [kHostStreams][kVelBins] worlds per velocity-sort key in the current sub-step
  uint32_t* bin_list = nullptr;     // [kVelBins][n] range-relative world indices filed under each key
  int sm_count = 148;
  uint64_t dbg_toi_worlds = 0, dbg_toi_total = 0;
  cudaStream_t pstream[kHostStreams] = {};   // world ranges of one pipeline call run on streams of their own
  cudaEvent_t pev[kHostStreams] = {}, pev_begin = nullptr;
};

namespace {

size_t smem_bytes(const BLCD_PENV* h, int block) { return (size_t)kSceneBytes + (size_t)h->scene.hot_words * block * sizeof(float); }
// the velocity kernel keeps only velocity and mass rows on chip, the position kernel only position and mass rows (blcd_pipeline.cuh)
size_t vel_smem_bytes(const BLCD_PENV* h) { return (size_t)kSceneBytes + (size_t)pipe_vel_hot_words(h->scene) * kPipeBlock * sizeof(float); }

template <typename F>
int launch_sized(BLCD_PENV* h, F f) {
  switch (h->block) {
#if BLCD_PROFILE_ID == 0
    case 64: return f(std::integral_constant<int, 64>());
#endif
    case 128: return f(std::integral_constant<int, 128>());
#if BLCD_PROFILE_ID == 0
    case 160: return f(std::integral_constant<int, 160>());
    case 192: return f(std::integral_constant<int, 192>());
#endif
    case 224: return f(std::integral_constant<int, 224>());
    case 256: return f(std::integral_constant<int, 256>());
#if BLCD_PROFILE_ID == 0   // the large profile is only ever launched with 256 threads (or 128 as the shared-memory fallback)
    case 320: return f(std::integral_constant<int, 320>());
    case 384: return f(std::integral_constant<int, 384>());
    case 448: return f(std::integral_constant<int, 448>());
    case 512: return f(std::integral_constant<int, 512>());
#endif
    default: return fail("unsupported block size");
  }
}

template <typename K>
int set_smem_attr(K kernel, size_t bytes, int blocks_per_sm = 1) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // shared memory for `blocks_per_sm` resident blocks; the rest of the 256 KB on-chip array serves as L1 for the
  // thread-local constraint records
  int pct = (int)((bytes + 1024) * blocks_per_sm * 100 / (228 * 1024)) + 1;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct));
  return 0;
}

int begin_timing(BLCD_PENV* h, cudaStream_t st) {
  if (h->timing) CK(cudaEventRecord(h->ev0, st));
  return 0;
}
int end_timing(BLCD_PENV* h, cudaStream_t st) {
  if (h->timing) { CK(cudaEventRecord(h->ev1, st)); h->last_ms = -2.0f; }
  return 0;
}

}  // namespace

extern "C" {

int BLCD_P(create)(const blcd_spec* spec_host, int64_t n_worlds, int device, uint64_t seed, int64_t world_offset, BLCD_PENV** out) {
  if (!spec_host || !out || n_worlds <= 0) return fail("blcd_create: bad arguments");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("blcd_create: no such CUDA device");
  CK(cudaSetDevice(device));
  BLCD_PENV* h = new BLCD_PENV();
  int maxm = host::default_manifold_slots(spec_host->n_bodies);
  if (const char* e = getenv("BLCD_MAX_MANIFOLDS")) maxm = atoi(e);
  if (maxm < 1 || maxm > kMaxSlots) { delete h; return fail("BLCD_MAX_MANIFOLDS out of range for this profile"); }
  const char* err = host::build_scene(h->scene, *spec_host, maxm);
  if (err) { delete h; return fail(std::string("blcd_create: ") + err); }
  if (const char* e = getenv("BLCD_ALIGN")) h->scene.align_mode = atoi(e);
  h->n = n_worlds; h->device = device; h->seed = seed; h->world_offset = world_offset;
  int smem_max = 0;
  CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  // One block per SM (the register file is the occupancy limit) whose warps walk the solver phases together
  // (Sim::phase_align).  A launch therefore takes waves x (time of one block), waves = ceil(blocks / SMs).
  //  * articulated scenes: 256 threads.  Below 253 registers per thread the joint records spill, and a block's time grows
  //    almost as fast as its size (Urchin at four full waves: 384 threads take 1.43x as long as 256 for 1.5x the worlds,
  //    i.e. +5 %; Luxo +4 %; LuxoCube -8 %), so larger blocks are not worth their partly filled last wave.
  //  * joint-free scenes (balls, boxes): a block's time grows slowly with its size (Bounce2: 448 threads 1.26), so the size
  //    with the smallest waves x time estimate for this world count wins -- 65 536 Bounce2 worlds fit ONE wave of
  //    448-thread blocks instead of two of 256 (+33 %), 262 144 take 4 waves instead of 7 (+17 %).
  int sm_count = 148;
  CK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  h->sm_count = sm_count;
  //  * few worlds (less than one 256-thread block per SM): smaller blocks, so that more SMs have one -- a block's time
  //    shrinks only a little with its size (128 threads: 0.83 of 256), but 4 096 worlds then use 64 SMs instead of 16.
  //  * between one half and one full wave of 256-thread blocks: the smallest block size that still gives every SM at most one
  //    block, so that all SMs work -- 32 768 worlds (262 144 split over 8 GPUs) are 128 blocks of 256 threads, i.e. 20 idle
  //    SMs, but 147 blocks of 224.
  h->block = 256;
  if (n_worlds > (int64_t)sm_count * 128 && n_worlds < (int64_t)sm_count * 256 && BLCD_PROFILE_ID == 0) {
    for (int b : {160, 192, 224})
      if ((n_worlds + b - 1) / b <= sm_count) { h->block = b; break; }
  } else if (n_worlds <= (int64_t)sm_count * 128) {
#if BLCD_PROFILE_ID == 0
    h->block = n_worlds <= (int64_t)sm_count * 64 ? 64 : 128;
#else
    h->block = 128;
#endif
  } else if (h->scene.nj == 0 || BLCD_PROFILE_ID == 0) {
    // relative time of one block by size: joint-free scenes grow slowly; articulated ones almost in proportion (their joint
    // records spill above 256 threads), so a larger block only pays when it saves a whole wave -- 65 536 Urchin worlds are
    // 1.73 waves of 256-thread blocks but ONE wave of 448 (measured 15.7 M env-steps/s against 14.2 M; LuxoCube 9.4 against 8.7)
    const int sizes[] = {256, 320, 384, 448, 512};
    const double rel_free[] = {1.00, 1.09, 1.175, 1.26, 1.35}, rel_joint[] = {1.00, 1.27, 1.43, 1.81, 1.91};
    const double* rel = h->scene.nj == 0 ? rel_free : rel_joint;
    double best = 0.0;
    for (int i = 0; i < 5; ++i) {
      if (smem_bytes(h, sizes[i]) > (size_t)smem_max) continue;
      const int64_t blocks = (n_worlds + sizes[i] - 1) / sizes[i], waves = (blocks + sm_count - 1) / sm_count;
      const double cost = (double)waves * rel[i];
      if (best == 0.0 || cost < best * 0.995) { best = cost; h->block = sizes[i]; }   // near-ties keep the smaller block
    }
  }
#if BLCD_PROFILE_ID == 1
  // large scenes: shared memory allows 256 or 224 threads; the smaller block wins when it fills the last wave (65 536 worlds:
  // 1.73 waves of 256 but 1.98 of 224 -- CrabCube 1.27 -> 1.38 M env-steps/s, SpiderCube 2.86 -> 3.54 M)
  if (n_worlds > (int64_t)sm_count * 128) {
    const int sizes[] = {256, 224};
    const double rel[] = {1.00, 0.90};
    double best = 0.0;
    for (int i = 0; i < 2; ++i) {
      if (smem_bytes(h, sizes[i]) > (size_t)smem_max) continue;
      const int64_t blocks = (n_worlds + sizes[i] - 1) / sizes[i], waves = (blocks + sm_count - 1) / sm_count;
      const double cost = (double)waves * rel[i];
      if (best == 0.0 || cost < best * 0.995) { best = cost; h->block = sizes[i]; }
    }
  }
#endif
  if (const char* e = getenv("BLCD_BLOCK")) h->block = atoi(e);
  // Which device path steps this handle's worlds (both give the same results up to FMA-contraction-level round-off; each is
  // deterministic and independent of how the worlds are sharded):
  //  * the phase pipeline (blcd_pipeline.cuh) once there are enough worlds to fill the GPU in every phase -- measured on a
  //    B200 (Urchin, M env-steps/s, pipeline against fused): 28.7 / 16.4 at 262 144 worlds, 21.7 / 14.9 at 131 072, 17.3 / 15.2 at
  //    98 304, 15.2 / 13.9 at 81 920, but 13.9 / 15.7 at 65 536 (its five kernels per sub-step each need their own wave of
  //    worlds); joint-free scenes cross over later (Bounce2: 46.7 / 51.5 at 98 304, 87 / 52 at 262 144);
  //  * the fused one-thread-per-world kernel below that, and for single environments.
  //    the crab-class scenes of the large build already at 65 536 (CrabCube 2.38 / 1.91, SpiderCube 4.95 / 4.79; eight
  //    world ranges, pipeline_run);
  h->pipeline = n_worlds >= (BLCD_PROFILE_ID == 1 ? 65536 : (h->scene.nj > 0 ? 77824 : 106496));
  if (const char* e = getenv("BLCD_PIPELINE")) h->pipeline = atoi(e) != 0;
  {
#if BLCD_PROFILE_ID == 0
    const int sizes[] = {512, 448, 384, 320, 256, 224, 192, 160, 128, 64};
#else
    const int sizes[] = {256, 224, 128};
#endif
    bool known = false;
    for (int b : sizes) known |= (b == h->block);
    if (!known) h->block = 256;
    for (int b : sizes)   // the largest supported size not above the request whose working set fits shared memory
      if (b <= h->block && smem_bytes(h, b) <= (size_t)smem_max) { h->block = b; break; }
  }
  if (smem_bytes(h, h->block) > (size_t)smem_max) { delete h; return fail("scene working set does not fit shared memory"); }
  // everything below allocates device resources: a failure (e.g. out of memory for too many worlds) releases what was
  // already allocated, so that a caller can retry with fewer worlds without leaking
  int rc = [&]() -> int {
    CK(cudaMalloc(&h->scene_dev, sizeof(DScene)));
    CK(cudaMemcpy(h->scene_dev, &h->scene, sizeof(DScene), cudaMemcpyHostToDevice));
    size_t bytes = (size_t)h->scene.state_words * (size_t)n_worlds * 4;
    CK(cudaMalloc(&h->state, bytes));
    CK(cudaMemset(h->state, 0, bytes));
    k_init<<<(unsigned)((n_worlds + 255) / 256), 256>>>(h->scene_dev, h->state, n_worlds);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    h->launches += 1;
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    return launch_sized(h, [&](auto B) {
      constexpr int BLOCK = decltype(B)::value;
      size_t sb = smem_bytes(h, BLOCK);
      if (set_smem_attr(k_reset<BLOCK>, sb)) return -1;
      if (set_smem_attr(k_set_bodies<BLOCK>, sb)) return -1;
      if (set_smem_attr(k_step<BLOCK>, sb)) return -1;
      if (set_smem_attr(k_rollout<BLOCK>, sb)) return -1;
      if (set_smem_attr(k_observe<BLOCK>, sb)) return -1;
      return 0;
    });
  }();
  if (rc) {
    cudaGetLastError();          // clear the sticky-free error state before the frees
    BLCD_P(destroy)(h);          // frees whatever was allocated (null members are skipped); keeps the recorded message
    return rc;
  }
  *out = h;
  return 0;
}

// Re-key the handle's random streams: world w becomes global world world_offset + w of stream family `seed`.  Only the key
// changes (and every world's draw counter restarts at zero, as in a fresh handle) -- the state buffer, the scene tables
// and every allocation stay; the next blcd_reset samples from the new streams, so a re-keyed handle produces exactly
// what a handle created with (seed, world_offset) would.  Lets a collector walk one allocation over a long dataset.
int BLCD_P(rekey)(BLCD_PENV* h, uint64_t seed, int64_t world_offset) {
  if (!h) return fail("blcd_rekey: null handle");
  if (h->host_submitted != h->host_completed) return fail("blcd_rekey: host steps are still in flight");
  CK(cudaSetDevice(h->device));
  // draw counters are one [word][world] row of the state buffer
  CK(cudaMemset(h->state + (size_t)(h->scene.off_misc + 3) * (size_t)h->n, 0, (size_t)h->n * 4));
  h->seed = seed;
  h->world_offset = world_offset;
  return 0;
}

int BLCD_P(destroy)(BLCD_PENV* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->dbg_toi_total) fprintf(stderr, "[blcd] worlds reaching the TOI kernel: %.4f of world-sub-steps\n", (double)h->dbg_toi_worlds / (double)h->dbg_toi_total);
  for (auto& st : h->hstream) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }   // nothing in flight past here
  cudaFree(h->scene_dev);
  cudaFree(h->state);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->h_act) cudaFreeHost(h->h_act);
  if (h->h_fs) cudaFreeHost(h->h_fs);
  if (h->h_bits) cudaFreeHost(h->h_bits);
  if (h->h_done) cudaFreeHost(h->h_done);
  cudaFree(h->d_act); cudaFree(h->d_fs); cudaFree(h->d_bits); cudaFree(h->d_done);
  for (auto& s : h->pstream) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
  for (auto& e : h->pev) if (e) cudaEventDestroy(e);
  if (h->pev_begin) cudaEventDestroy(h->pev_begin);
  cudaFree(h->scratch); cudaFree(h->toi_count); cudaFree(h->toi_list); cudaFree(h->pos_next); cudaFree(h->bin_count); cudaFree(h->bin_list);
  for (auto& r : h->pinned_user) cudaHostUnregister(const_cast<void*>(r.first));
  for (auto& ring : h->hev) for (auto& e : ring) if (e) cudaEventDestroy(e);
  delete h;
  return 0;
}

int BLCD_P(reset)(BLCD_PENV* h, const int64_t* idx_dev, int64_t n, const float* full_state_dev, uint64_t stream) {
  if (!h) return fail("blcd_reset: null handle");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int64_t cnt = idx_dev ? n : h->n;
  if (cnt <= 0) return 0;
  int rc = launch_sized(h, [&](auto B) {
    constexpr int BLOCK = decltype(B)::value;
    k_reset<BLOCK><<<(unsigned)((cnt + BLOCK - 1) / BLOCK), BLOCK, smem_bytes(h, BLOCK), st>>>(h->scene_dev, h->state, h->n, h->seed, h->world_offset,
                                                                                              idx_dev, cnt, full_state_dev);
    return 0;
  });
  if (rc) return rc;
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int BLCD_P(set_bodies)(BLCD_PENV* h, const float* bodies_dev, const uint32_t* variant_dev, uint64_t stream) {
  if (!h || !bodies_dev) return fail("blcd_set_bodies: bad arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc = launch_sized(h, [&](auto B) {
    constexpr int BLOCK = decltype(B)::value;
    k_set_bodies<BLOCK><<<(unsigned)((h->n + BLOCK - 1) / BLOCK), BLOCK, smem_bytes(h, BLOCK), st>>>(h->scene_dev, h->state, h->n, h->seed, h->world_offset,
                                                                                                     bodies_dev, variant_dev);
    return 0;
  });
  if (rc) return rc;
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int BLCD_P(get_bodies)(BLCD_PENV* h, float* bodies_dev, uint64_t stream) {
  if (!h || !bodies_dev) return fail("blcd_get_bodies: bad arguments");
  CK(cudaSetDevice(h->device));
  k_get_bodies<<<(unsigned)((h->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->scene_dev, h->state, h->n, bodies_dev);
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int BLCD_P(check_finite)(BLCD_PENV* h, uint8_t* invalid_dev, int64_t* n_invalid_host) {
  if (!h || !n_invalid_host) return fail("blcd_check_finite: bad arguments");
  CK(cudaSetDevice(h->device));
  unsigned long long* cnt = nullptr;
  CK(cudaMalloc(&cnt, sizeof(unsigned long long)));
  CK(cudaMemset(cnt, 0, sizeof(unsigned long long)));
  k_check_finite<<<(unsigned)((h->n + 255) / 256), 256>>>(h->scene_dev, h->state, h->n, invalid_dev, cnt);
  cudaError_t e = cudaGetLastError();
  unsigned long long host = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&host, cnt, sizeof(host), cudaMemcpyDeviceToHost);
  cudaFree(cnt);
  if (e != cudaSuccess) return fail(std::string("blcd_check_finite: ") + cudaGetErrorString(e));
  *n_invalid_host = (int64_t)host;
  h->launches += 1;
  return 0;
}

int BLCD_P(get_poses)(BLCD_PENV* h, float* poses_dev, uint32_t* variant_dev, uint64_t stream) {
  if (!h || !poses_dev) return fail("blcd_get_poses: bad arguments");
  CK(cudaSetDevice(h->device));
  k_get_poses<<<(unsigned)((h->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->scene_dev, h->state, h->n, poses_dev, variant_dev);
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

// ---- phase pipeline: host side ---------------------------------------------------------------------------------------
static int pipeline_prepare(BLCD_PENV* h) {
  if (h->scratch) return 0;
  CK(cudaMalloc(&h->scratch, (size_t)h->scene.scratch_words * (size_t)h->n * 4));
  CK(cudaMemset(h->scratch, 0, (size_t)h->scene.scratch_words * (size_t)h->n * 4));   // the sweep-count hints start at zero
  CK(cudaMalloc(&h->toi_count, kHostStreams * sizeof(uint32_t)));
  CK(cudaMemset(h->toi_count, 0, kHostStreams * sizeof(uint32_t)));
  CK(cudaMalloc(&h->toi_list, (size_t)h->n * sizeof(uint32_t)));
  CK(cudaMalloc(&h->bin_count, kHostStreams * kAllBins * sizeof(uint32_t)));
  CK(cudaMemset(h->bin_count, 0, kHostStreams * kAllBins * sizeof(uint32_t)));
  CK(cudaMalloc(&h->bin_list, (size_t)kAllBins * (size_t)h->n * sizeof(uint32_t)));
  CK(cudaMalloc(&h->pos_next, kHostStreams * sizeof(unsigned long long)));
  CK(cudaMemset(h->pos_next, 0, kHostStreams * sizeof(unsigned long long)));
  const size_t sb = smem_bytes(h, kPipeBlock);
  // shared memory for the blocks each kernel is compiled to keep resident; the rest of the SM's array is L1, which the
  // velocity / position kernels need for their thread-local contact records
  const size_t sbv = vel_smem_bytes(h);
  const int vel_fit = (int)((227 * 1024) / (sbv + 1024));
  if (set_smem_attr(k_pipe_pre, sb, kPreBlocks) || set_smem_attr(k_pipe_vel, sbv, vel_fit < kVelBlocks ? vel_fit : kVelBlocks) ||
      set_smem_attr(k_pipe_pos, sbv, vel_fit < kPosBlocks ? vel_fit : kPosBlocks) ||
      set_smem_attr(k_pipe_post, sb, kPostBlocks) || set_smem_attr(k_pipe_toi, smem_bytes(h, kToiBlock), 16))
    return -1;
  for (int r = 0; r < kHostStreams; ++r) {
    if (!h->pstream[r]) CK(cudaStreamCreateWithFlags(&h->pstream[r], cudaStreamNonBlocking));
    if (!h->pev[r]) CK(cudaEventCreateWithFlags(&h->pev[r], cudaEventDisableTiming));
  }
  if (!h->pev_begin) CK(cudaEventCreateWithFlags(&h->pev_begin, cudaEventDisableTiming));
  return 0;
}

// T env steps of worlds [w0, w1) on stream st: per sub-step five launches.  mode 0: blcd_step semantics (actions_dev or the
// device RNG, T == 1); mode 1: blcd_rollout semantics (obs_t / a_t recorded before step t).  `slot`: which TOI counter to
// use (ranges running concurrently on different streams need their own).
// one sub-step (five launches) of worlds [w0, w1) on stream st
static void pipeline_substep(BLCD_PENV* h, const float* actions_dev, int mode, int T, int t, int s, OutPtrs out, cudaStream_t st, int64_t w0, int64_t w1, int slot) {
  const unsigned blocks = (unsigned)((w1 - w0 + kPipeBlock - 1) / kPipeBlock);
  const size_t sb = smem_bytes(h, kPipeBlock);
  uint32_t* cnt = h->toi_count + slot;
  uint32_t* list = h->toi_list + w0;
  uint32_t* bins = h->bin_count + slot * kAllBins;
  // the position kernel is persistent: as many blocks as can be resident, never more than there are worlds for
  static const int pos_resident = getenv("BLCD_POS_BLOCKS") ? atoi(getenv("BLCD_POS_BLOCKS")) : kPosBlocks;
  const int pos_fit = (int)((227 * 1024) / (vel_smem_bytes(h) + 1024));
  const int pos_cap = pos_fit < kPosBlocks ? (pos_fit < 1 ? 1 : pos_fit) : kPosBlocks;
  unsigned pos_blocks = (unsigned)(h->sm_count * (pos_resident < 1 ? 1 : (pos_resident > pos_cap ? pos_cap : pos_resident)));
  if (pos_blocks > blocks) pos_blocks = blocks;
  static const int pos_sort = getenv("BLCD_POS_SORT") ? atoi(getenv("BLCD_POS_SORT")) : 2;   // 0 world order, 1 by previous sweep count, 2 by contact count
  static const int pos_refill = getenv("BLCD_POS_REFILL") ? atoi(getenv("BLCD_POS_REFILL")) : 16;
  {
    {
      if (mode == 1 && s == 0 && out.lcd_bits) {
        // obs_t's frame: rendered from the state as it is before step t by the body-major render kernel (3x the lanes of the
        // per-thread renderer inside k_pipe_pre); k_pipe_pre then records full_state and the action only
        if (!h->render_attr_set) {
          cudaFuncSetAttribute(k_render_poses, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSceneBytes + sizeof(BodyPx) * kMaxBodies * kRenderMaxFrames + sizeof(RowInk) * kRenderThreads));
          cudaFuncSetAttribute(k_render_bodies, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSceneBytes + sizeof(RowInk) * kR2Frames * 256));
          h->render_attr_set = true;
        }
        const int64_t cnt_w = w1 - w0;
        k_render_bodies<<<(unsigned)((cnt_w + kR2Frames - 1) / kR2Frames), kR2Threads, (size_t)kSceneBytes + sizeof(RowInk) * kR2Frames * (size_t)h->scene.lcd_h, st>>>(
            h->scene_dev, nullptr, nullptr, h->state, h->n, w0, cnt_w, h->scene.lcd_w, h->scene.lcd_h, out.lcd_bits + (size_t)w0 * T * h->scene.lcd_h * row_words(h->scene.lcd_w),
            0, h->scene.lcd_w, row_words(h->scene.lcd_w), 0, T, t);
        h->launches += 1;
        out.lcd_bits = nullptr;
      }
      k_pipe_pre<<<blocks, kPipeBlock, sb, st>>>(h->scene_dev, h->state, h->scratch, h->n, h->seed, h->world_offset, actions_dev, mode, T, t, s == 0 ? 1 : 0, out,
                                                 w0, w1, cnt, h->pos_next + slot, bins, h->bin_list, pos_sort == 1);
      static const bool vel_sort = !(getenv("BLCD_VEL_SORT") && atoi(getenv("BLCD_VEL_SORT")) == 0);
      k_pipe_vel<<<blocks, kPipeBlock, vel_smem_bytes(h), st>>>(h->scene_dev, h->state, h->scratch, h->n, w0, w1, bins, vel_sort ? h->bin_list : nullptr);
      k_pipe_pos<<<pos_blocks, kPipeBlock, vel_smem_bytes(h), st>>>(h->scene_dev, h->state, h->scratch, h->n, w0, w1, h->pos_next + slot, pos_refill, bins, pos_sort ? h->bin_list : nullptr, pos_sort);
      k_pipe_post<<<blocks, kPipeBlock, sb, st>>>(h->scene_dev, h->state, h->scratch, h->n, h->seed, h->world_offset, w0, w1, cnt, list, bins);
      static const bool dbg_toi = getenv("BLCD_DEBUG_TOI") != nullptr;   // diagnostic: share of worlds that reach the TOI kernel
      if (dbg_toi) {
        uint32_t c = 0;
        cudaMemcpyAsync(&c, cnt, 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        h->dbg_toi_worlds += c; h->dbg_toi_total += (uint64_t)(w1 - w0);
      }
      k_pipe_toi<<<(unsigned)((w1 - w0 + kToiBlock - 1) / kToiBlock), kToiBlock, smem_bytes(h, kToiBlock), st>>>(h->scene_dev, h->state, h->scratch, h->n, h->seed, h->world_offset, w0, cnt, list);
      h->launches += 5;
    }
  }
}

// T env steps of worlds [w0, w1), ordered after / before the work on stream st.  The worlds are split into up to kHostStreams
// ranges, each on a stream of its own with its own TOI list / sort bins / work counter, and the launches of the ranges are
// interleaved: while one range sits in a latency-bound phase (narrow phase, TOI, the tail of any kernel) another range's
// velocity kernel fills the SMs.  `slot0`: first counter slot to use (blcd_step_host's concurrent sub-ranges pass theirs).
static int pipeline_run(BLCD_PENV* h, const float* actions_dev, int mode, int T, OutPtrs out, cudaStream_t st, int64_t w0, int64_t w1, int slot0, int max_ranges = kHostStreams) {
  if (pipeline_prepare(h)) return -1;
  if (w1 <= w0) return 0;
  // measured (Urchin): 131 072 worlds 19.6 M env-steps/s as one range, 20.9 M as two, 21.7 M as four; 262 144 worlds 25.6 / 26.2 / 25.1 M
  static const int want = getenv("BLCD_PIPE_RANGES") ? atoi(getenv("BLCD_PIPE_RANGES")) : 0;
  // crab-class scenes (large build; two or three resident blocks per SM in every kernel): CrabCube 65 536 worlds 2.11 / 2.21 / 2.25 /
  // 2.36 M as 1 / 2 / 4 / 8 ranges, 131 072 worlds 2.53 / 2.80 / 2.88 / 2.89 M; SpiderCube indifferent from 4 on
  int R = want > 0 ? want : (BLCD_PROFILE_ID == 1 ? 8 : ((w1 - w0) < 196608 ? 4 : 2));
  R = R > max_ranges ? max_ranges : R;
  const int64_t gran = 4 * kPipeBlock;
  static const int64_t min_range = getenv("BLCD_PIPE_MIN_RANGE") ? atoll(getenv("BLCD_PIPE_MIN_RANGE")) : (BLCD_PROFILE_ID == 1 ? 8192 : 24576);
  while (R > 1 && (w1 - w0) / R < min_range) --R;    // small ranges cannot fill the GPU: the interleaving only pays with enough worlds each
  if (R == 1) {
    for (int t = 0; t < T; ++t)
      for (int s = 0; s < h->scene.nsub; ++s) pipeline_substep(h, actions_dev, mode, T, t, s, out, st, w0, w1, slot0);
    CK(cudaGetLastError());
    return 0;
  }
  CK(cudaEventRecord(h->pev_begin, st));
  for (int r = 0; r < R; ++r) CK(cudaStreamWaitEvent(h->pstream[r], h->pev_begin, 0));
  for (int t = 0; t < T; ++t)
    for (int s = 0; s < h->scene.nsub; ++s)
      for (int r = 0; r < R; ++r) {
        const int64_t a = w0 + ((w1 - w0) * r / R) / gran * gran, b = r + 1 == R ? w1 : w0 + ((w1 - w0) * (r + 1) / R) / gran * gran;
        if (b > a) pipeline_substep(h, actions_dev, mode, T, t, s, out, h->pstream[r], a, b, slot0 + r);
      }
  CK(cudaGetLastError());
  for (int r = 0; r < R; ++r) {
    CK(cudaEventRecord(h->pev[r], h->pstream[r]));
    CK(cudaStreamWaitEvent(st, h->pev[r], 0));
  }
  return 0;
}

static int step_range(BLCD_PENV* h, const float* actions_dev, int n_steps, OutPtrs out, cudaStream_t st, int64_t w_begin, int64_t w_end, int slot = 0) {
  if (h->pipeline) {
    OutPtrs o = out;
    OutPtrs act_only = {nullptr, nullptr, nullptr, nullptr, nullptr, out.actions};
    for (int i = 0; i < n_steps; ++i)
      if (pipeline_run(h, actions_dev, 0, 1, act_only, st, w_begin, w_end, slot, w_begin == 0 && w_end == h->n ? kHostStreams : 1)) return -1;
    o.actions = nullptr;
    if (o.lcd_bits) {   // the frames by the body-major render kernel, straight from the state; k_observe then only packs full_state / proprio / done
      if (!h->render_attr_set) {
        CK(cudaFuncSetAttribute(k_render_poses, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSceneBytes + sizeof(BodyPx) * kMaxBodies * kRenderMaxFrames + sizeof(RowInk) * kRenderThreads)));
        CK(cudaFuncSetAttribute(k_render_bodies, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSceneBytes + sizeof(RowInk) * kR2Frames * 256)));
        h->render_attr_set = true;
      }
      const int64_t cnt_w = w_end - w_begin;
      const int lw = row_words(h->scene.lcd_w);
      k_render_bodies<<<(unsigned)((cnt_w + kR2Frames - 1) / kR2Frames), kR2Threads, (size_t)kSceneBytes + sizeof(RowInk) * kR2Frames * (size_t)h->scene.lcd_h, st>>>(
          h->scene_dev, nullptr, nullptr, h->state, h->n, w_begin, cnt_w, h->scene.lcd_w, h->scene.lcd_h, o.lcd_bits + (size_t)w_begin * h->scene.lcd_h * lw,
          0, h->scene.lcd_w, lw, 0, 1, 0);
      CK(cudaGetLastError());
      h->launches += 1;
      o.lcd_bits = nullptr;
    }
    if (o.full_state || o.proprio || o.lcd_bool || o.done) {
      int rc = launch_sized(h, [&](auto B) {
        constexpr int BLOCK = decltype(B)::value;
        k_observe<BLOCK><<<(unsigned)((w_end - w_begin + BLOCK - 1) / BLOCK), BLOCK, smem_bytes(h, BLOCK), st>>>(h->scene_dev, h->state, h->n, o, w_begin, w_end);
        return 0;
      });
      if (rc) return rc;
      CK(cudaGetLastError());
      h->launches += 1;
    }
    return 0;
  }
  int rc = launch_sized(h, [&](auto B) {
    constexpr int BLOCK = decltype(B)::value;
    k_step<BLOCK><<<(unsigned)((w_end - w_begin + BLOCK - 1) / BLOCK), BLOCK, smem_bytes(h, BLOCK), st>>>(h->scene_dev, h->state, h->n, h->seed, h->world_offset,
                                                                                                          actions_dev, n_steps, out, w_begin, w_end);
    return 0;
  });
  if (rc) return rc;
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

static int step_impl(BLCD_PENV* h, const float* actions_dev, int n_steps, OutPtrs out, cudaStream_t st) {
  if (begin_timing(h, st)) return -1;
  if (step_range(h, actions_dev, n_steps, out, st, 0, h->n)) return -1;
  return end_timing(h, st);
}

int BLCD_P(step)(BLCD_PENV* h, const float* actions_dev, float* actions_out_dev, uint64_t stream) {
  if (!h) return fail("blcd_step: null handle");
  CK(cudaSetDevice(h->device));
  OutPtrs out = {nullptr, nullptr, nullptr, nullptr, nullptr, actions_out_dev};
  return step_impl(h, actions_dev, 1, out, (cudaStream_t)stream);
}

int BLCD_P(step_observe)(BLCD_PENV* h, const float* actions_dev, float* actions_out_dev, float* full_state_dev, float* proprio_dev,
                      uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev, uint8_t* done_dev, uint64_t stream) {
  if (!h) return fail("blcd_step_observe: null handle");
  CK(cudaSetDevice(h->device));
  OutPtrs out = {full_state_dev, proprio_dev, lcd_bits_dev, lcd_bool_dev, done_dev, actions_out_dev};
  return step_impl(h, actions_dev, 1, out, (cudaStream_t)stream);
}

int BLCD_P(observe)(BLCD_PENV* h, float* full_state_dev, float* proprio_dev, uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev, uint8_t* done_dev,
                 uint64_t stream) {
  if (!h) return fail("blcd_observe: null handle");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  OutPtrs out = {full_state_dev, proprio_dev, lcd_bits_dev, lcd_bool_dev, done_dev, nullptr};
  int rc = launch_sized(h, [&](auto B) {
    constexpr int BLOCK = decltype(B)::value;
    k_observe<BLOCK><<<(unsigned)((h->n + BLOCK - 1) / BLOCK), BLOCK, smem_bytes(h, BLOCK), st>>>(h->scene_dev, h->state, h->n, out, 0, h->n);
    return 0;
  });
  if (rc) return rc;
  CK(cudaGetLastError());
  h->launches += 1;
  return 0;
}

int BLCD_P(rollout)(BLCD_PENV* h, int32_t T, float* full_state_dev, uint32_t* lcd_bits_dev, float* actions_dev, uint64_t stream) {
  if (!h || T <= 0) return fail("blcd_rollout: bad arguments");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  OutPtrs out = {full_state_dev, nullptr, lcd_bits_dev, nullptr, nullptr, actions_dev};
  if (begin_timing(h, st)) return -1;
  if (h->pipeline) {
    if (pipeline_run(h, nullptr, 1, T, out, st, 0, h->n, 0)) return -1;
    return end_timing(h, st);
  }
  int rc = launch_sized(h, [&](auto B) {
    constexpr int BLOCK = decltype(B)::value;
    k_rollout<BLOCK><<<(unsigned)((h->n + BLOCK - 1) / BLOCK), BLOCK, smem_bytes(h, BLOCK), st>>>(h->scene_dev, h->state, h->n, h->seed, h->world_offset, T, out);
    return 0;
  });
  if (rc) return rc;
  CK(cudaGetLastError());
  if (end_timing(h, st)) return -1;
  h->launches += 1;
  return 0;
}

// pin a caller-owned host buffer once (page-locks it in place) so that later copies are direct DMA transfers; returns
// false if the driver refuses (then the call falls back to the handle's own pinned staging buffers)
static bool is_pinned(const BLCD_PENV* h, const void* p, size_t bytes) {
  const char* q = static_cast<const char*>(p);
  for (auto& r : h->pinned_user) {
    const char* base = static_cast<const char*>(r.first);
    if (q >= base && q + bytes <= base + r.second) return true;
  }
  return false;
}

// page-lock caller memory on the caller's explicit request (blcd_pin_host); it stays registered until blcd_unpin_host or
// blcd_destroy, and the caller keeps it alive for that long
static bool pin_user_buffer(BLCD_PENV* h, const void* p, size_t bytes) {
  if (!p) return false;
  if (is_pinned(h, p, bytes)) return true;
  if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  h->pinned_user.emplace_back(p, bytes);
  return true;
}

// Enqueue one env step through host buffers on the handle's pipeline streams.  The worlds are split into sub-ranges, one
// stream each: the copies of one range overlap the kernel of the next, and the ragged tail of one kernel (blocks finish
// at different times) is filled by the first blocks of the next -- also across consecutive steps, because range c of
// step t+1 only waits for range c of step t.  The streams are ordinary blocking streams, so everything stays ordered
// after earlier work on the legacy default stream (reset, device-side steps) and before later work on it.
// src / dst: page-locked memory the copies read from / write to (the caller's buffers, or the handle's staging area).
static int host_step_submit(BLCD_PENV* h, const float* act_src, float* fs_dst, uint32_t* bits_dst, uint8_t* done_dst, int chunks) {
  const DScene& sc = h->scene;
  const int64_t gran = 4 * (int64_t)h->block;   // range boundaries stay multiples of the block size
  if (chunks < 1) chunks = 1;
  if (chunks > kHostStreams) chunks = kHostStreams;
  while (chunks > 1 && h->n / chunks < gran) --chunks;
  if (h->host_chunks && h->host_chunks != chunks && h->host_submitted != h->host_completed)
    return fail("blcd_step_host: the number of pipeline ranges cannot change while steps are in flight");
  h->host_chunks = chunks;
  for (int c = 0; c < chunks; ++c)
    if (!h->hstream[c]) CK(cudaStreamCreate(&h->hstream[c]));
  const int lw = row_words(sc.lcd_w);
  const int slot = (int)(h->host_submitted % kHostDepth);
  if (chunks == 1 && begin_timing(h, h->hstream[0])) return -1;
  for (int c = 0; c < chunks; ++c) {
    const int64_t w0 = (h->n * c / chunks) / gran * gran, w1 = c + 1 == chunks ? h->n : (h->n * (c + 1) / chunks) / gran * gran;
    cudaStream_t st = h->hstream[c];
    if (w1 > w0) {
      const size_t cnt = (size_t)(w1 - w0);
      if (act_src) CK(cudaMemcpyAsync(h->d_act + w0 * sc.A, act_src + w0 * sc.A, cnt * sc.A * 4, cudaMemcpyHostToDevice, st));
      OutPtrs out = {fs_dst ? h->d_fs : nullptr, nullptr, bits_dst ? h->d_bits : nullptr, nullptr, done_dst ? h->d_done : nullptr, nullptr};
      if (step_range(h, act_src ? h->d_act : nullptr, 1, out, st, w0, w1, c)) return -1;
      if (fs_dst) CK(cudaMemcpyAsync(fs_dst + w0 * sc.S, h->d_fs + w0 * sc.S, cnt * sc.S * 4, cudaMemcpyDeviceToHost, st));
      if (bits_dst) CK(cudaMemcpyAsync(bits_dst + w0 * sc.lcd_h * lw, h->d_bits + w0 * sc.lcd_h * lw, cnt * sc.lcd_h * lw * 4, cudaMemcpyDeviceToHost, st));
      if (done_dst) CK(cudaMemcpyAsync(done_dst + w0, h->d_done + w0, cnt, cudaMemcpyDeviceToHost, st));
    }
    if (!h->hev[slot][c]) CK(cudaEventCreateWithFlags(&h->hev[slot][c], cudaEventDisableTiming));
    CK(cudaEventRecord(h->hev[slot][c], st));
  }
  if (chunks == 1 && end_timing(h, h->hstream[0])) return -1;
  h->host_submitted += 1;
  return 0;
}

// block until at most `keep` submitted host steps are still in flight
static int host_step_wait(BLCD_PENV* h, int keep) {
  if (keep < 0) keep = 0;
  while (h->host_submitted - h->host_completed > (uint64_t)keep) {
    const int slot = (int)(h->host_completed % kHostDepth);
    for (int c = 0; c < h->host_chunks; ++c) CK(cudaEventSynchronize(h->hev[slot][c]));
    h->host_completed += 1;
  }
  return 0;
}

static int host_step_buffers(BLCD_PENV* h, size_t* na, size_t* nf, size_t* nb, size_t* nd) {
  const DScene& sc = h->scene;
  *na = (size_t)h->n * sc.A * 4; *nf = (size_t)h->n * sc.S * 4; *nb = (size_t)h->n * sc.lcd_h * row_words(sc.lcd_w) * 4; *nd = (size_t)h->n;
  if (!h->d_act) {
    CK(cudaMalloc(&h->d_act, *na)); CK(cudaMalloc(&h->d_fs, *nf)); CK(cudaMalloc(&h->d_bits, *nb)); CK(cudaMalloc(&h->d_done, *nd));
    CK(cudaMallocHost(&h->h_act, *na)); CK(cudaMallocHost(&h->h_fs, *nf)); CK(cudaMallocHost(&h->h_bits, *nb)); CK(cudaMallocHost(&h->h_done, *nd));
  }
  return 0;
}

static int host_chunks_default(BLCD_PENV* h) {
  if (h->host_chunks_env < 0) {   // read the override once per handle
    const char* e = getenv("BLCD_HOST_CHUNKS");
    h->host_chunks_env = e ? atoi(e) : 0;
  }
  if (h->host_chunks_env > 0) return h->host_chunks_env;
  if (h->timing) return 1;
  if (h->pipeline) return h->n >= 4 * 98304 ? 4 : (h->n >= 2 * 98304 ? 2 : 1);   // every range must still fill the GPU phase by phase (each is split further by pipeline_run only when it is the whole handle)
  // fused kernel: one block per SM, so a handle of at most one wave of blocks is stepped as ONE range (splitting it would round
  // every range up to whole blocks and spill into a second wave: 32 768 worlds in 8 ranges are 152 blocks of 224 on 148 SMs);
  // larger handles overlap the copies of one range with the kernel of the next
  const int64_t blocks = (h->n + h->block - 1) / h->block;
  if (blocks <= h->sm_count) return 1;
  return blocks <= 2 * (int64_t)h->sm_count ? 2 : kHostStreams;
}

int BLCD_P(step_host)(BLCD_PENV* h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host) {
  if (!h) return fail("blcd_step_host: null handle");
  CK(cudaSetDevice(h->device));
  size_t na, nf, nb, nd;
  if (host_step_buffers(h, &na, &nf, &nb, &nd)) return -1;
  if (host_step_wait(h, 0)) return -1;   // earlier asynchronous steps first: the device-side staging area is shared
  // Buffers the caller page-locked with blcd_pin_host are copied to / from directly (DMA); anything else is staged
  // through the handle's own pinned memory.  The call never page-locks caller memory itself: a Python caller passes
  // fresh numpy arrays every step, and a registration that outlives its array would go stale when the allocator
  // unmaps and re-uses the address.
  const bool pa = actions_host && is_pinned(h, actions_host, na), pf = full_state_host && is_pinned(h, full_state_host, nf);
  const bool pb = lcd_bits_host && is_pinned(h, lcd_bits_host, nb), pd = done_host && is_pinned(h, done_host, nd);
  if (actions_host && !pa) memcpy(h->h_act, actions_host, na);
  if (host_step_submit(h, actions_host ? (pa ? actions_host : h->h_act) : nullptr, full_state_host ? (pf ? full_state_host : h->h_fs) : nullptr,
                       lcd_bits_host ? (pb ? lcd_bits_host : h->h_bits) : nullptr, done_host ? (pd ? done_host : h->h_done) : nullptr,
                       host_chunks_default(h)))
    return -1;
  if (host_step_wait(h, 0)) return -1;
  if (full_state_host && !pf) memcpy(full_state_host, h->h_fs, nf);
  if (lcd_bits_host && !pb) memcpy(lcd_bits_host, h->h_bits, nb);
  if (done_host && !pd) memcpy(done_host, h->h_done, nd);
  return 0;
}

int BLCD_P(pin_host)(BLCD_PENV* h, const void* buf_host, int64_t bytes) {
  if (!h || !buf_host || bytes <= 0) return fail("blcd_pin_host: bad arguments");
  CK(cudaSetDevice(h->device));
  if (!pin_user_buffer(h, buf_host, (size_t)bytes)) return fail("blcd_pin_host: the driver refused to page-lock this buffer");
  return 0;
}

int BLCD_P(unpin_host)(BLCD_PENV* h, const void* buf_host) {
  if (!h || !buf_host) return fail("blcd_unpin_host: bad arguments");
  CK(cudaSetDevice(h->device));
  for (size_t i = 0; i < h->pinned_user.size(); ++i)
    if (h->pinned_user[i].first == buf_host) {
      if (host_step_wait(h, 0)) return -1;     // nothing may still be copying from / into it
      CK(cudaHostUnregister(const_cast<void*>(buf_host)));
      h->pinned_user.erase(h->pinned_user.begin() + i);
      return 0;
    }
  return fail("blcd_unpin_host: this address was not registered with blcd_pin_host");
}

int BLCD_P(step_host_async)(BLCD_PENV* h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host) {
  if (!h) return fail("blcd_step_host_async: null handle");
  CK(cudaSetDevice(h->device));
  size_t na, nf, nb, nd;
  if (host_step_buffers(h, &na, &nf, &nb, &nd)) return -1;
  if ((actions_host && !is_pinned(h, actions_host, na)) || (full_state_host && !is_pinned(h, full_state_host, nf)) ||
      (lcd_bits_host && !is_pinned(h, lcd_bits_host, nb)) || (done_host && !is_pinned(h, done_host, nd)))
    return fail("blcd_step_host_async: every buffer must lie inside memory page-locked with blcd_pin_host");
  // the device-side staging area is single-buffered per world range: range c of this step runs after range c of the
  // previous step on the same stream, so its copies are ordered; bound the queue by the event ring
  if (host_step_wait(h, kHostDepth - 1)) return -1;
  return host_step_submit(h, actions_host, full_state_host, lcd_bits_host, done_host, host_chunks_default(h));
}

int BLCD_P(step_host_wait)(BLCD_PENV* h, int32_t keep_in_flight) {
  if (!h) return fail("blcd_step_host_wait: null handle");
  CK(cudaSetDevice(h->device));
  return host_step_wait(h, keep_in_flight);
}

int BLCD_P(render_poses)(BLCD_PENV* h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, uint32_t* lcd_bits_dev, uint64_t stream) {
  return BLCD_P(render_poses_sized)(h, poses_dev, variant_dev, n, 0, 0, lcd_bits_dev, stream);
}

int BLCD_P(render_poses_sized)(BLCD_PENV* h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, int32_t lcd_w, int32_t lcd_h,
                            uint32_t* lcd_bits_dev, uint64_t stream) {
  if (!h || !poses_dev || !lcd_bits_dev || n < 0) return fail("blcd_render_poses: bad arguments");
  if (n == 0) return 0;
  if (lcd_w <= 0) lcd_w = h->scene.lcd_w;
  if (lcd_h <= 0) lcd_h = h->scene.lcd_h;
  if (lcd_w > 1024) return fail("blcd_render_poses: frame width out of range (<= 1024)");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (lcd_h > kRenderThreads) return fail("blcd_render_poses: frame height out of range");
  static const bool old_kernel = getenv("BLCD_RENDER_ROWS") && atoi(getenv("BLCD_RENDER_ROWS")) != 0;   // (frame, row)-per-lane kernel, kept for comparison
  const int fpb = render_frames_per_block(lcd_h);
  const size_t rsm_bytes = (size_t)kSceneBytes + sizeof(BodyPx) * kMaxBodies * (size_t)fpb + sizeof(RowInk) * (size_t)kRenderThreads;
  if (!h->render_attr_set) {
    CK(cudaFuncSetAttribute(k_render_poses, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSceneBytes + sizeof(BodyPx) * kMaxBodies * kRenderMaxFrames + sizeof(RowInk) * kRenderThreads)));
    CK(cudaFuncSetAttribute(k_render_bodies, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kSceneBytes + sizeof(RowInk) * kR2Frames * 256)));
    h->render_attr_set = true;
  }
  if (begin_timing(h, st)) return -1;
  const int row_stride = row_words(lcd_w);
  for (int x_off = 0; x_off < lcd_w; x_off += kRowBits) {   // one launch per column window of kRowBits pixels
    const int win_w = lcd_w - x_off < kRowBits ? lcd_w - x_off : kRowBits;
    if (old_kernel)
      k_render_poses<<<(unsigned)((n + fpb - 1) / fpb), kRenderThreads, rsm_bytes, st>>>(h->scene_dev, poses_dev, variant_dev, n, lcd_w, lcd_h, lcd_bits_dev,
                                                                                         x_off, win_w, row_stride, x_off / 32);
    else
      k_render_bodies<<<(unsigned)((n + kR2Frames - 1) / kR2Frames), kR2Threads, (size_t)kSceneBytes + sizeof(RowInk) * kR2Frames * (size_t)lcd_h, st>>>(
          h->scene_dev, poses_dev, variant_dev, nullptr, 0, 0, n, lcd_w, lcd_h, lcd_bits_dev, x_off, win_w, row_stride, x_off / 32, 1, 0);
    CK(cudaGetLastError());
    h->launches += 1;
  }
  if (end_timing(h, st)) return -1;
  return 0;
}

int64_t BLCD_P(state_bytes)(BLCD_PENV* h) { return h ? (int64_t)h->scene.state_words * h->n * 4 : -1; }

int BLCD_P(save_state)(BLCD_PENV* h, void* buf_dev, uint64_t stream) {
  if (!h || !buf_dev) return fail("blcd_save_state: bad arguments");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(buf_dev, h->state, (size_t)BLCD_P(state_bytes)(h), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int BLCD_P(load_state)(BLCD_PENV* h, const void* buf_dev, uint64_t stream) {
  if (!h || !buf_dev) return fail("blcd_load_state: bad arguments");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->state, buf_dev, (size_t)BLCD_P(state_bytes)(h), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int64_t BLCD_P(num_worlds)(BLCD_PENV* h) { return h ? h->n : -1; }
int64_t BLCD_P(kernel_launches)(BLCD_PENV* h) { return h ? h->launches : -1; }

int BLCD_P(enable_timing)(BLCD_PENV* h, int on) {
  if (!h) return fail("blcd_enable_timing: null handle");
  h->timing = on != 0;
  return 0;
}

int BLCD_P(last_step_ms)(BLCD_PENV* h, float* ms_out) {
  if (!h || !ms_out) return fail("blcd_last_step_ms: bad arguments");
  if (!h->timing || h->last_ms == -1.0f) return fail("blcd_last_step_ms: timing not enabled or nothing timed yet");
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms_out, h->ev0, h->ev1));
  return 0;
}

int BLCD_P(get_counters)(BLCD_PENV* h, uint32_t* counters_dev, uint64_t stream) {
  if (!h || !counters_dev) return fail("blcd_get_counters: bad arguments");
  CK(cudaSetDevice(h->device));
  // counters are BLCD_N_COUNTERS consecutive [word][world] rows: transpose with a strided 2-D copy
  for (int k = 0; k < BLCD_N_COUNTERS; ++k)
    CK(cudaMemcpy2DAsync(counters_dev + k, BLCD_N_COUNTERS * 4, h->state + (size_t)(h->scene.off_cnt + k) * h->n, 4, 4, (size_t)h->n,
                         cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int BLCD_P(scene_info)(BLCD_PENV* h, int32_t* out16) {
  if (!h || !out16) return fail("blcd_scene_info: bad arguments");
  const DScene& sc = h->scene;
  int32_t v[16] = {sc.nb, sc.nj, sc.nw, sc.np, sc.S, sc.P, sc.A, sc.lcd_w, sc.lcd_h, sc.maxm, sc.state_words, sc.hot_words, h->block,
                   (int32_t)smem_bytes(h, h->block), BLCD_PROFILE_ID, h->pipeline};
  memcpy(out16, v, sizeof(v));
  return 0;
}

}  // extern "C"
