// blcd_profile.h -- compile-time scene-size profile of the simulation source.
//
// The per-world working set (body rows in shared memory, constraint records in thread-local memory, bit masks over
// bodies / pairs / frame columns in registers) is sized by compile-time limits.  The library carries the SAME sources
// compiled twice:
//   small  (default)             <= 8 bodies, <= 7 joints, <= 64 collidable pairs, frames <= 32 px wide
//                                every scene of envs.py:17-110 -- the headline path; masks fit one register (pair)
//   large  (-DBLCD_PROFILE_LARGE) <= 18 bodies, <= 17 joints, <= 128 pairs, frames <= 64 px wide
//                                Crab / CrabCube / SpiderCube (envs.py:116-137, lcd_base=32)
// Each build lives in its own namespace (and exports its C entry points under its own prefix, boxlcd_b200.cu);
// blcd_dispatch.cpp owns the public blcd_* symbols and routes a handle to the profile that blcd_create picked.
#pragma once
#include <stdint.h>

#if defined(BLCD_PROFILE_LARGE)
#define BLCD_NS blcd_large
#define BLCD_PROFILE_ID 1
#define BLCD_PROFILE_NAME "large"
#else
#define BLCD_NS blcd_small
#define BLCD_PROFILE_ID 0
#define BLCD_PROFILE_NAME "small"
#endif

#if defined(__CUDACC__)
#define BLCD_PHD __host__ __device__ __forceinline__
#else
#define BLCD_PHD inline
#endif

namespace BLCD_NS {

#if defined(BLCD_PROFILE_LARGE)
constexpr int kMaxBodies = 18, kMaxJoints = 17, kMaxPairs = 128, kMaxSlots = 32;
typedef uint64_t RowMask;   // one frame row, bit x = pixel x
#else
constexpr int kMaxBodies = 8, kMaxJoints = 7, kMaxPairs = 64, kMaxSlots = 16;
typedef uint32_t RowMask;
#endif
constexpr int kMaxObs = 4 * kMaxBodies;
constexpr int kRowBits = 8 * (int)sizeof(RowMask);

// set of candidate-pair indices (0 .. kMaxPairs-1)
struct PairMask {
  static constexpr int W = (kMaxPairs + 63) / 64;
  uint64_t w[W];
  BLCD_PHD static PairMask none() { PairMask m; for (int i = 0; i < W; ++i) m.w[i] = 0ull; return m; }
  BLCD_PHD static PairMask all() { PairMask m; for (int i = 0; i < W; ++i) m.w[i] = ~0ull; return m; }
  BLCD_PHD static PairMask bit(int p) { PairMask m = none(); m.set(p); return m; }
  BLCD_PHD void set(int p) { if (W == 1) w[0] |= 1ull << p; else w[p >> 6] |= 1ull << (p & 63); }
  BLCD_PHD void clear(int p) { if (W == 1) w[0] &= ~(1ull << p); else w[p >> 6] &= ~(1ull << (p & 63)); }
  BLCD_PHD bool test(int p) const { return W == 1 ? ((w[0] >> p) & 1ull) != 0 : ((w[p >> 6] >> (p & 63)) & 1ull) != 0; }
};

}  // namespace BLCD_NS
