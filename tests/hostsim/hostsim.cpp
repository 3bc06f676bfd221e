// hostsim.cpp -- TEST TOOL: compiles the CUDA path's per-world simulation source (boxlcd_b200/csrc/blcd_world.cuh)
// for the host, so its logic can be compared bit-for-bit with the CPU oracle in a container that has no GPU.
// It is never loaded by the product; the product path is libboxlcd_b200.so on a GPU and nothing else.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <new>
#include "blcd_pipeline.cuh"

using namespace BLCD_NS;

struct HostSim {
  DScene scene;
  std::vector<uint32_t> state;
  std::vector<float> hot;
  std::vector<uint32_t> scratch;   // phase pipeline: [word][world], like the device buffer
  int64_t n, offset;
  uint64_t seed;
};

extern "C" {

void* hostsim_new(const blcd_spec* spec, int64_t n, uint64_t seed, int64_t offset, int maxm) {
  HostSim* h = new HostSim();
  if (maxm <= 0) maxm = host::default_manifold_slots(spec->n_bodies);  // same default as blcd_create
  if (host::build_scene(h->scene, *spec, maxm)) { delete h; return nullptr; }
  h->n = n; h->seed = seed; h->offset = offset;
  h->state.assign((size_t)h->scene.state_words * n, 0u);
  h->hot.assign((size_t)h->scene.hot_words, 0.0f);
  h->scratch.assign((size_t)h->scene.scratch_words * n, 0xDEADBEEFu);
  for (int64_t w = 0; w < n; ++w)
    for (int s = 0; s < h->scene.maxm; ++s) h->state[(size_t)(h->scene.off_slots + kSlotWords * s) * n + w] = kSlotFree;
  return h;
}
void hostsim_free(void* p) { delete (HostSim*)p; }

#define SIM(w) Sim<1> sim(h->scene, h->hot.data(), h->state.data(), h->n, (w)); sim.load(h->seed, h->offset + (w))

void hostsim_reset(void* p, const float* full_state) {
  HostSim* h = (HostSim*)p;
  for (int64_t w = 0; w < h->n; ++w) {
    SIM(w);
    sim.reset(full_state ? full_state + w * h->scene.S : nullptr);
    sim.store();
  }
}

void hostsim_set_bodies(void* p, const float* bodies, const uint32_t* variants) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  for (int64_t w = 0; w < h->n; ++w) {
    SIM(w);
    sim.variant = variants ? (variants[w] & kBodyMask) : 0u;
    float pose[kMaxBodies][3];
    const float* src = bodies + w * sc.nb * BLCD_BODY_STATE;
    for (int b = 0; b < sc.nb; ++b) { pose[b][0] = src[b * 6]; pose[b][1] = src[b * 6 + 1]; pose[b][2] = src[b * 6 + 2]; }
    sim.build_fresh(pose);
    for (int b = 0; b < sc.nb; ++b) { sim.v[b] = mk(src[b * 6 + 3], src[b * 6 + 4]); sim.w[b] = src[b * 6 + 5]; }
    sim.store();
  }
}

void hostsim_get_bodies(void* p, float* bodies) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  for (int64_t w = 0; w < h->n; ++w) {
    SIM(w);
    for (int b = 0; b < sc.nb; ++b) {
      float* d = bodies + (w * sc.nb + b) * BLCD_BODY_STATE;
      d[0] = sim.xf[b].p.x; d[1] = sim.xf[b].p.y; d[2] = sim.a[b]; d[3] = sim.v[b].x; d[4] = sim.v[b].y; d[5] = sim.w[b];
    }
  }
}

static void write_obs(const Sim<1>& sim, const DScene& sc, float* fs_out, uint32_t* bits_out) {
  if (fs_out)
    for (int b = 0; b < sc.nb; ++b) {
      float ob[4];
      sim.obs_body(b, ob);
      for (int k = 0; k < 4; ++k) fs_out[sc.body[b].obs[k]] = ob[k];
    }
  if (bits_out) {
    BodyPx bp[kMaxBodies];
    for (int b = 0; b < sc.nb; ++b) body_px(bp[b], sim.bshape(b), sim.xf[b].p.x, sim.xf[b].p.y, sim.xf[b].q.s, sim.xf[b].q.c, sc.world_w, sc.lcd_w);
    const int lw = row_words(sc.lcd_w);
    for (int R = 0; R < sc.lcd_h; ++R) {
      RowMask ink = 0u;
      for (int b = 0; b < sc.nb; ++b) ink |= body_px_row(bp[b], sc.lcd_h - 1 - R, sc.lcd_w, sc.lcd_h, sc.rules);
      for (int k = 0; k < lw; ++k) bits_out[R * lw + k] = row_word(row_bits_from_ink(ink, sc.lcd_w), k);
    }
  }
}

void hostsim_step(void* p, const float* actions, float* actions_out) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  for (int64_t w = 0; w < h->n; ++w) {
    SIM(w);
    float act[BLCD_MAX_OBS];
    if (actions) memcpy(act, actions + w * sc.A, sizeof(float) * sc.A);
    else sim.draw_action(act);
    if (actions_out) memcpy(actions_out + w * sc.A, act, sizeof(float) * sc.A);
    sim.env_step(act);
    sim.store();
  }
}

// ---- the phase pipeline (blcd_pipeline.cuh) on the host: every phase runs in a FRESH, poisoned Sim object, so anything a
// phase needs must really travel through the state / scratch buffers, exactly as between two kernels on the device -------
namespace {
struct Phase {
  alignas(64) unsigned char buf[sizeof(Sim<1>)];
  Sim<1>* sim;
  Phase(HostSim* h, int64_t w, bool load) {
    memset(buf, 0xAB, sizeof(buf));
    for (auto& x : h->hot) { uint32_t poison = 0xABABABABu; memcpy(&x, &poison, 4); }
    sim = new (buf) Sim<1>(h->scene, h->hot.data(), h->state.data(), h->n, w);
    sim->attach_scratch(h->scratch.data(), w);
    if (load) sim->load(h->seed, h->offset + w);
  }
};

void pipeline_env_step(HostSim* h, int64_t w, const float* act) {
  for (int s = 0; s < h->scene.nsub; ++s) {
    { Phase p(h, w, true); pipe_pre(*p.sim, s == 0, act); }
    { Phase p(h, w, false); pipe_vel(*p.sim); }
    // the device kernel is persistent: one Sim object comes through pipe_pos_begin once per world it fetches, so the begin
    // must not depend on what an earlier begin left in the object -- entered twice here to catch that
    { Phase p(h, w, false); pipe_pos_begin(*p.sim); pipe_pos(*p.sim); }
    bool need;
    { Phase p(h, w, true); need = pipe_post(*p.sim); }
    if (need) { Phase p(h, w, true); pipe_toi(*p.sim); }
  }
}
}  // namespace

void hostsim_step_pipeline(void* p, const float* actions, float* actions_out) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  for (int64_t w = 0; w < h->n; ++w) {
    float act[BLCD_MAX_OBS];
    if (actions) memcpy(act, actions + w * sc.A, sizeof(float) * sc.A);
    else { SIM(w); sim.draw_action(act); sim.store(); }
    if (actions_out) memcpy(actions_out + w * sc.A, act, sizeof(float) * sc.A);
    pipeline_env_step(h, w, act);
  }
}

void hostsim_rollout_pipeline(void* p, int T, float* full_state, uint32_t* bits, float* actions) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  for (int64_t w = 0; w < h->n; ++w) {
    for (int t = 0; t < T; ++t) {
      int64_t row = w * T + t;
      float act[BLCD_MAX_OBS];
      {
        SIM(w);
        write_obs(sim, sc, full_state ? full_state + row * sc.S : nullptr, bits ? bits + row * sc.lcd_h * row_words(sc.lcd_w) : nullptr);
        sim.draw_action(act);
        sim.store();
      }
      if (actions) memcpy(actions + row * sc.A, act, sizeof(float) * sc.A);
      pipeline_env_step(h, w, act);
    }
  }
}

void hostsim_observe(void* p, float* full_state, uint32_t* bits) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  for (int64_t w = 0; w < h->n; ++w) {
    SIM(w);
    write_obs(sim, sc, full_state ? full_state + w * sc.S : nullptr, bits ? bits + w * sc.lcd_h * row_words(sc.lcd_w) : nullptr);
  }
}

void hostsim_rollout(void* p, int T, float* full_state, uint32_t* bits, float* actions) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  for (int64_t w = 0; w < h->n; ++w) {
    SIM(w);
    float act[BLCD_MAX_OBS];
    for (int t = 0; t < T; ++t) {
      int64_t row = w * T + t;
      write_obs(sim, sc, full_state ? full_state + row * sc.S : nullptr, bits ? bits + row * sc.lcd_h * row_words(sc.lcd_w) : nullptr);
      sim.draw_action(act);
      if (actions) memcpy(actions + row * sc.A, act, sizeof(float) * sc.A);
      sim.env_step(act);
    }
    sim.store();
  }
}

// polygon_row_t (the converged-warp scanline rules of k_render_bodies) against polygon_row on random integer polygons:
// convex hull-like quads, degenerate ones (collinear / repeated vertices after truncation), 3..8 vertices, rows inside and
// outside, both rule sets, column windows.  Returns the number of mismatching (polygon, row) cases.
int64_t hostsim_polygon_row_check(int64_t trials, uint64_t seed) {
  uint64_t s = seed * 2654435761ull + 88172645463325252ull;
  auto rnd = [&](int m) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (int)(s % (uint64_t)m); };
  int64_t bad = 0;
  for (int64_t t = 0; t < trials; ++t) {
    PolyPx P;
    const int kind = rnd(8);
    P.n = kind < 5 ? 4 : 3 + rnd(6);
    const int cx = rnd(40) - 4, cy = rnd(24) - 4, span = 1 + rnd(kind == 0 ? 3 : 12);
    if (kind < 3) {            // rotated box-like quad: centre + two half-diagonals
      int ax = rnd(2 * span + 1) - span, ay = rnd(2 * span + 1) - span, bx = -ay + rnd(3) - 1, by = ax + rnd(3) - 1;
      if (kind == 2) { bx /= 3; by /= 3; }
      int vx[4] = {cx - ax - bx, cx + ax - bx, cx + ax + bx, cx - ax + bx}, vy[4] = {cy - ay - by, cy + ay - by, cy + ay + by, cy - ay + by};
      for (int i = 0; i < 4; ++i) { P.x[i] = vx[i]; P.y[i] = vy[i]; }
    } else {
      for (int i = 0; i < P.n; ++i) { P.x[i] = cx + rnd(2 * span + 1) - span; P.y[i] = cy + rnd(2 * span + 1) - span; }
      if (rnd(4) == 0 && P.n > 1) { P.x[P.n - 1] = P.x[0]; P.y[P.n - 1] = P.y[0]; }
    }
    for (int i = P.n; i < BLCD_MAX_VERTS; ++i) { P.x[i] = 12345; P.y[i] = -777; }
    int ylo = P.y[0], yhi = P.y[0];
    for (int i = 1; i < P.n; ++i) { ylo = P.y[i] < ylo ? P.y[i] : ylo; yhi = P.y[i] > yhi ? P.y[i] : yhi; }
    PolySlopes S;
    polygon_slopes(S, P);
    const int h = 8 + rnd(3) * 8, w = rnd(2) ? 32 : 24, rules = rnd(2), x_off = rnd(4) == 0 ? 8 * rnd(3) : 0;
    for (int y = ylo - 1; y <= yhi + 1; ++y) {
      RowMask a = polygon_row(P, y, w, h, rules, ylo, yhi, x_off), b = polygon_row_fast(P, S, y, w, h, rules, ylo, yhi, x_off);
      bad += a != b;
    }
  }
  return bad;
}

void hostsim_counters(void* p, uint32_t* out) {
  HostSim* h = (HostSim*)p;
  for (int64_t w = 0; w < h->n; ++w)
    for (int k = 0; k < BLCD_N_COUNTERS; ++k) out[w * BLCD_N_COUNTERS + k] = h->state[(size_t)(h->scene.off_cnt + k) * h->n + w];
}

void hostsim_render_poses(void* p, const float* poses, const uint32_t* variants, int64_t n, int lcd_w, int lcd_h, uint32_t* bits) {
  HostSim* h = (HostSim*)p;
  const DScene& sc = h->scene;
  if (lcd_w <= 0) lcd_w = sc.lcd_w;
  if (lcd_h <= 0) lcd_h = sc.lcd_h;
  const int lw = row_words(lcd_w);
  for (int64_t w = 0; w < n; ++w)
    for (int R = 0; R < lcd_h; ++R)
      for (int x_off = 0; x_off < lcd_w; x_off += kRowBits) {   // column windows of one RowMask, like blcd_render_poses_sized
        const int win_w = lcd_w - x_off < kRowBits ? lcd_w - x_off : kRowBits;
        RowMask ink = 0u;
        uint32_t variant = variants ? variants[w] : 0u;
        for (int b = 0; b < sc.nb; ++b) {
          const float* q = poses + (w * sc.nb + b) * 4;
          ink |= body_row(sc.body[b].shape[(variant >> b) & 1u], q[0], q[1], q[2], q[3], lcd_h - 1 - R, sc.world_w, lcd_w, lcd_h, sc.rules, x_off, win_w);
        }
        for (int k = 0; k < row_words(win_w); ++k) bits[(w * lcd_h + R) * lw + x_off / 32 + k] = row_word(row_bits_from_ink(ink, win_w), k);
      }
}

}  // extern "C"
