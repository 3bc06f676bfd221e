// blcd_pipeline.cuh -- the PHASE PIPELINE: one sub-step of b2World::Step as five kernels over all worlds instead of one
// fused thread per world.
//
// Why: the fused kernel (k_step / k_rollout) keeps a world on one thread for the whole rollout, so (i) its register
// budget is the sum of all phases (253 registers -> 8 warps per SM), (ii) all warps of a block must walk the phases in
// lock step to share the instruction cache, so every phase lasts as long as the slowest of 256 worlds (28 % of the time
// is barrier wait), (iii) lanes whose world has nothing to do in a phase idle (12.6 of 32 lanes active).  HBM is idle
// (0.03 % of peak).  The pipeline spends that idle bandwidth: between phases the per-world records go through a scratch
// area in HBM ([word][world], coalesced), so every phase is its own small kernel with its own register budget and
// occupancy, no block-level barriers, and lists of only the worlds that need the rare phases:
//     k_pipe_pre    [obs_t, a_t, SetMotorSpeed on the first sub-step]  Collide + islands + constraint setup + warm start
//     k_pipe_vel    180 velocity iterations + StoreImpulses + position integration   (lean: rows, joint / contact records)
//     k_pipe_pos    position iterations
//     k_pipe_post   write-back, sleeping, broad phase; decides whether SolveTOI has anything to do (Sim::toi_needed)
//     k_pipe_toi    SolveTOI for the listed worlds, end of the sub-step
// The phase bodies are the SAME functions the fused path runs (Sim::solve_setup / solve_velocity / ...), in the same order,
// so both paths give the same results; they are written once here for device and host (tests/hostsim runs them through
// a host scratch buffer, a fresh Sim per phase, and diffs against the CPU oracle bit for bit).
#pragma once
#include "blcd_world.cuh"

namespace BLCD_NS {

// phase 1.  Caller: sim.load().  `first`: this is the first sub-step of an env step -> apply the action.
template <int S>
BLCD_HD void pipe_pre(Sim<S>& sim, bool first, const float* action) {
  if (first) sim.env_step_begin(action);
  sim.substep_collide();
  sim.solve_setup(sim.scene().dt, sim.inv_dt0 * sim.scene().dt);
  sim.islDone = 0u;
  sim.x_rows_out(0, 8);
  sim.x_jr_out(0, kHotJoint);
  sim.x_cr_out();
  sim.x_misc_out();
  sim.store_after_setup();
}

// shared-memory words per thread of the velocity and position kernels' lean columns
BLCD_HD int pipe_vel_hot_words(const DScene& sc) { return 5 * (sc.nb + 1); }

// phase 2.  No load(): everything comes from the scratch area; impulses go to the manifold slots / scratch.
template <int S>
BLCD_HD void pipe_vel(Sim<S>& sim) {
  sim.x_misc_in();
  sim.oM = sim.oP;   // lean column: [v 3 x (nb+1)][invMass invI 2 x (nb+1)], positions stay in the scratch rows (pipe_vel_hot_words)
  sim.x_rows_in_lean();
  sim.x_jr_in(0, kHotJoint);
  sim.x_cr_in();
  sim.template solve_velocity<true>(sim.scene().dt);
  sim.solve_integrate_x(sim.scene().dt);
  sim.x_jr_out(J_IX, 4);
}

// phase 3.  Needs the shape variants (local centres, radii) besides the scratch records.
template <int S>
BLCD_HD void pipe_pos_begin(Sim<S>& sim) {
  sim.load_variant();
  sim.x_misc_in();
  sim.oP = sim.oV; sim.oM = sim.oV + 3 * (sim.scene().nb + 1);   // (absolute: the persistent kernel comes through here once per world)
  // lean column: [c a 3 x (nb+1)][invMass invI 2 x (nb+1)], no velocity rows in this kernel
  sim.x_rows_in(3, 5);
  sim.x_cr_pk_in();
  for (int k = 0; k < sim.nc; ++k) sim.pos_cache_manifold(k);   // manifold data is read once per world, not once per sweep
  sim.x_jr_in(J_MM, 1);
  sim.x_jr_in(J_PK, 2);       // packed limit state + J_REF
  sim.cnt[BLCD_CNT_POS_ITERS] = 0u;
  sim.islDone = 0u;
}
// word of the scratch area holding how many position sweeps the world needed last time (scheduling hint, see k_pipe_pos)
BLCD_HD int pipe_prev_sweeps_word(const DScene& sc) { return sc.x_c0 - 1; }

template <int S>
BLCD_HD void pipe_pos_end(Sim<S>& sim, int sweeps = 0) {
  sim.x.u(pipe_prev_sweeps_word(sim.scene())) = (uint32_t)sweeps;
  sim.x_rows_out(3, 3);
  sim.x.u(sim.scene().x_misc + 1) = sim.islDone;
  sim.g.u(sim.scene().off_cnt + BLCD_CNT_POS_ITERS) += sim.cnt[BLCD_CNT_POS_ITERS];
}
template <int S>
BLCD_HD void pipe_pos(Sim<S>& sim) {
  pipe_pos_begin(sim);
  for (int it = 0; it < sim.scene().pos_iters; ++it)
    if (sim.template solve_position_sweep<true>()) break;
  pipe_pos_end(sim);
}

// phase 4.  Caller: sim.load() (poses / transforms before the solve = c0, a0, xf1).  Returns true if the world goes on to
// the TOI kernel; otherwise the sub-step is complete.
template <int S>
BLCD_HD bool pipe_post(Sim<S>& sim) {
  sim.x_misc_in();
  sim.x_rows_in(0, 6);
  sim.x_jr_in(J_IX, 4);
  sim.x_jr_in(J_PK, 1);
  sim.solve_finish(sim.scene().dt);
  bool need = !(sim.scene().flags & BLCD_FLAG_NO_TOI) && sim.toi_needed();
  if (need) {
    for (int b = 0; b < sim.scene().nb; ++b) {
      sim.x.f(sim.scene().x_c0 + 3 * b) = sim.c0[b].x;
      sim.x.f(sim.scene().x_c0 + 3 * b + 1) = sim.c0[b].y;
      sim.x.f(sim.scene().x_c0 + 3 * b + 2) = sim.a0[b];
    }
  } else {
    sim.substep_end();
  }
  sim.store();
  return need;
}

// phase 5.  Caller: sim.load().
template <int S>
BLCD_HD void pipe_toi(Sim<S>& sim) {
  for (int b = 0; b < sim.scene().nb; ++b) {
    sim.c0[b] = mk(sim.x.f(sim.scene().x_c0 + 3 * b), sim.x.f(sim.scene().x_c0 + 3 * b + 1));
    sim.a0[b] = sim.x.f(sim.scene().x_c0 + 3 * b + 2);
  }
  sim.solve_toi(sim.scene().dt);
  sim.substep_end();
  sim.store();
}

}  // namespace BLCD_NS
