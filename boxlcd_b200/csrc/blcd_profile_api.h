// blcd_profile_api.h -- the per-profile entry points behind the public C ABI (include/boxlcd_b200.h).
// Included with BLCD_P(name) = the prefixed symbol (blcd_small_<name> / blcd_large_<name>) and BLCD_PENV = the
// profile's handle type: once by each compilation of boxlcd_b200.cu (which defines them) and once per profile by
// blcd_dispatch.cpp (which calls them).  Signatures = the public ones with the profile's own handle type.
// No include guard on purpose.
#include <stdint.h>
#include "boxlcd_b200.h"

struct BLCD_PENV;

extern "C" {
int blcd_fail_msg(const char* msg);   // records the calling thread's error message (blcd_last_error), returns -1

int BLCD_P(create)(const blcd_spec* spec_host, int64_t n_worlds, int device, uint64_t seed, int64_t world_offset, BLCD_PENV** out);
int BLCD_P(destroy)(BLCD_PENV* h);
int BLCD_P(rekey)(BLCD_PENV* h, uint64_t seed, int64_t world_offset);
int BLCD_P(reset)(BLCD_PENV* h, const int64_t* idx_dev, int64_t n, const float* full_state_dev, uint64_t stream);
int BLCD_P(step)(BLCD_PENV* h, const float* actions_dev, float* actions_out_dev, uint64_t stream);
int BLCD_P(observe)(BLCD_PENV* h, float* full_state_dev, float* proprio_dev, uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev, uint8_t* done_dev,
                    uint64_t stream);
int BLCD_P(step_observe)(BLCD_PENV* h, const float* actions_dev, float* actions_out_dev, float* full_state_dev, float* proprio_dev,
                         uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev, uint8_t* done_dev, uint64_t stream);
int BLCD_P(rollout)(BLCD_PENV* h, int32_t T, float* full_state_dev, uint32_t* lcd_bits_dev, float* actions_dev, uint64_t stream);
int BLCD_P(step_host)(BLCD_PENV* h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host);
int BLCD_P(pin_host)(BLCD_PENV* h, const void* buf_host, int64_t bytes);
int BLCD_P(unpin_host)(BLCD_PENV* h, const void* buf_host);
int BLCD_P(step_host_async)(BLCD_PENV* h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host);
int BLCD_P(step_host_wait)(BLCD_PENV* h, int32_t keep_in_flight);
int BLCD_P(render_poses)(BLCD_PENV* h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, uint32_t* lcd_bits_dev, uint64_t stream);
int BLCD_P(render_poses_sized)(BLCD_PENV* h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, int32_t lcd_w, int32_t lcd_h,
                               uint32_t* lcd_bits_dev, uint64_t stream);
int BLCD_P(set_bodies)(BLCD_PENV* h, const float* bodies_dev, const uint32_t* variant_dev, uint64_t stream);
int BLCD_P(get_bodies)(BLCD_PENV* h, float* bodies_dev, uint64_t stream);
int BLCD_P(get_poses)(BLCD_PENV* h, float* poses_dev, uint32_t* variant_dev, uint64_t stream);
int64_t BLCD_P(state_bytes)(BLCD_PENV* h);
int BLCD_P(save_state)(BLCD_PENV* h, void* buf_dev, uint64_t stream);
int BLCD_P(load_state)(BLCD_PENV* h, const void* buf_dev, uint64_t stream);
int BLCD_P(check_finite)(BLCD_PENV* h, uint8_t* invalid_dev, int64_t* n_invalid_host);
int64_t BLCD_P(num_worlds)(BLCD_PENV* h);
int64_t BLCD_P(kernel_launches)(BLCD_PENV* h);
int BLCD_P(last_step_ms)(BLCD_PENV* h, float* ms_out);
int BLCD_P(enable_timing)(BLCD_PENV* h, int on);
int BLCD_P(get_counters)(BLCD_PENV* h, uint32_t* counters_dev, uint64_t stream);
int BLCD_P(scene_info)(BLCD_PENV* h, int32_t* out16);
}
