// blcd_dispatch.cpp -- the public C ABI of libboxlcd_b200.so (include/boxlcd_b200.h).
//
// The simulation source is compiled once per scene-size profile (blcd_profile.h); this file owns the blcd_* symbols,
// picks the profile in blcd_create from the spec, and forwards every later call on the handle to that build.
#include <cstdlib>
#include <cstring>
#include <string>
#include "boxlcd_b200.h"

#define BLCD_P(name) blcd_small_##name
#define BLCD_PENV blcd_small_env
#include "blcd_profile_api.h"
#undef BLCD_P
#undef BLCD_PENV
#define BLCD_P(name) blcd_large_##name
#define BLCD_PENV blcd_large_env
#include "blcd_profile_api.h"
#undef BLCD_P
#undef BLCD_PENV

struct blcd_env {
  int large;
  union { blcd_small_env* s; blcd_large_env* l; };
};

namespace {
thread_local std::string g_err;

// every scene of envs.py:17-110 fits the small profile (8 bodies x 4 walls + 28 body pairs <= 64 candidate pairs)
bool fits_small(const blcd_spec& sp) { return sp.n_bodies <= 8 && sp.n_joints <= 7 && sp.lcd_w <= 32; }
}  // namespace

// forward a call on handle h to the profile it was created with
#define BLCD_FWD(name, ...) (h->large ? blcd_large_##name(h->l, ##__VA_ARGS__) : blcd_small_##name(h->s, ##__VA_ARGS__))
#define BLCD_NEED(h, who) \
  if (!(h)) return blcd_fail_msg(who ": null handle")

extern "C" {

int blcd_fail_msg(const char* msg) { g_err = msg ? msg : ""; return -1; }
const char* blcd_last_error(void) { return g_err.c_str(); }
int blcd_version(void) { return 120; }

int blcd_create(const blcd_spec* spec_host, int64_t n_worlds, int device, uint64_t seed, int64_t world_offset, blcd_handle* out) {
  if (!spec_host || !out || n_worlds <= 0) return blcd_fail_msg("blcd_create: bad arguments");
  bool large = !fits_small(*spec_host);
  if (const char* e = getenv("BLCD_PROFILE")) {
    if (!strcmp(e, "large")) large = true;
    else if (strcmp(e, "small") && strcmp(e, "auto") && e[0]) return blcd_fail_msg("BLCD_PROFILE must be small, large or auto");
  }
  blcd_env* h = new blcd_env();
  h->large = large ? 1 : 0;
  int rc = large ? blcd_large_create(spec_host, n_worlds, device, seed, world_offset, &h->l)
                 : blcd_small_create(spec_host, n_worlds, device, seed, world_offset, &h->s);
  if (rc) { delete h; return rc; }
  *out = h;
  return 0;
}

int blcd_destroy(blcd_handle h) {
  if (!h) return 0;
  int rc = BLCD_FWD(destroy);
  delete h;
  return rc;
}

int blcd_rekey(blcd_handle h, uint64_t seed, int64_t world_offset) {
  BLCD_NEED(h, "blcd_rekey");
  return BLCD_FWD(rekey, seed, world_offset);
}
int blcd_reset(blcd_handle h, const int64_t* idx_dev, int64_t n, const float* full_state_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_reset");
  return BLCD_FWD(reset, idx_dev, n, full_state_dev, stream);
}
int blcd_step(blcd_handle h, const float* actions_dev, float* actions_out_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_step");
  return BLCD_FWD(step, actions_dev, actions_out_dev, stream);
}
int blcd_observe(blcd_handle h, float* full_state_dev, float* proprio_dev, uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev, uint8_t* done_dev,
                 uint64_t stream) {
  BLCD_NEED(h, "blcd_observe");
  return BLCD_FWD(observe, full_state_dev, proprio_dev, lcd_bits_dev, lcd_bool_dev, done_dev, stream);
}
int blcd_step_observe(blcd_handle h, const float* actions_dev, float* actions_out_dev, float* full_state_dev, float* proprio_dev,
                      uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev, uint8_t* done_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_step_observe");
  return BLCD_FWD(step_observe, actions_dev, actions_out_dev, full_state_dev, proprio_dev, lcd_bits_dev, lcd_bool_dev, done_dev, stream);
}
int blcd_rollout(blcd_handle h, int32_t T, float* full_state_dev, uint32_t* lcd_bits_dev, float* actions_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_rollout");
  return BLCD_FWD(rollout, T, full_state_dev, lcd_bits_dev, actions_dev, stream);
}
int blcd_step_host(blcd_handle h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host) {
  BLCD_NEED(h, "blcd_step_host");
  return BLCD_FWD(step_host, actions_host, full_state_host, lcd_bits_host, done_host);
}
int blcd_pin_host(blcd_handle h, const void* buf_host, int64_t bytes) {
  BLCD_NEED(h, "blcd_pin_host");
  return BLCD_FWD(pin_host, buf_host, bytes);
}
int blcd_unpin_host(blcd_handle h, const void* buf_host) {
  BLCD_NEED(h, "blcd_unpin_host");
  return BLCD_FWD(unpin_host, buf_host);
}
int blcd_step_host_async(blcd_handle h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host) {
  BLCD_NEED(h, "blcd_step_host_async");
  return BLCD_FWD(step_host_async, actions_host, full_state_host, lcd_bits_host, done_host);
}
int blcd_step_host_wait(blcd_handle h, int32_t keep_in_flight) {
  BLCD_NEED(h, "blcd_step_host_wait");
  return BLCD_FWD(step_host_wait, keep_in_flight);
}
int blcd_render_poses(blcd_handle h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, uint32_t* lcd_bits_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_render_poses");
  return BLCD_FWD(render_poses, poses_dev, variant_dev, n, lcd_bits_dev, stream);
}
int blcd_render_poses_sized(blcd_handle h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, int32_t lcd_w, int32_t lcd_h,
                            uint32_t* lcd_bits_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_render_poses");
  return BLCD_FWD(render_poses_sized, poses_dev, variant_dev, n, lcd_w, lcd_h, lcd_bits_dev, stream);
}
int blcd_set_bodies(blcd_handle h, const float* bodies_dev, const uint32_t* variant_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_set_bodies");
  return BLCD_FWD(set_bodies, bodies_dev, variant_dev, stream);
}
int blcd_get_bodies(blcd_handle h, float* bodies_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_get_bodies");
  return BLCD_FWD(get_bodies, bodies_dev, stream);
}
int blcd_get_poses(blcd_handle h, float* poses_dev, uint32_t* variant_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_get_poses");
  return BLCD_FWD(get_poses, poses_dev, variant_dev, stream);
}
int64_t blcd_state_bytes(blcd_handle h) { return h ? BLCD_FWD(state_bytes) : -1; }
int blcd_save_state(blcd_handle h, void* buf_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_save_state");
  return BLCD_FWD(save_state, buf_dev, stream);
}
int blcd_load_state(blcd_handle h, const void* buf_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_load_state");
  return BLCD_FWD(load_state, buf_dev, stream);
}
int blcd_check_finite(blcd_handle h, uint8_t* invalid_dev, int64_t* n_invalid_host) {
  BLCD_NEED(h, "blcd_check_finite");
  return BLCD_FWD(check_finite, invalid_dev, n_invalid_host);
}
int64_t blcd_num_worlds(blcd_handle h) { return h ? BLCD_FWD(num_worlds) : -1; }
int64_t blcd_kernel_launches(blcd_handle h) { return h ? BLCD_FWD(kernel_launches) : -1; }
int blcd_last_step_ms(blcd_handle h, float* ms_out) {
  BLCD_NEED(h, "blcd_last_step_ms");
  return BLCD_FWD(last_step_ms, ms_out);
}
int blcd_enable_timing(blcd_handle h, int on) {
  BLCD_NEED(h, "blcd_enable_timing");
  return BLCD_FWD(enable_timing, on);
}
int blcd_get_counters(blcd_handle h, uint32_t* counters_dev, uint64_t stream) {
  BLCD_NEED(h, "blcd_get_counters");
  return BLCD_FWD(get_counters, counters_dev, stream);
}
int blcd_scene_info(blcd_handle h, int32_t* out16) {
  BLCD_NEED(h, "blcd_scene_info");
  return BLCD_FWD(scene_info, out16);
}

}  // extern "C"
