/*
 * boxlcd_b200.h -- C ABI of libboxlcd_b200.so, the B200-native drop-in for boxLCD's hot path:
 * stepping N independent Box2D worlds in lockstep and rasterizing each into its binarized LCD frame.
 *
 * The reference has no FFI of its own (it is pure Python over pybox2d + Pillow); what this library
 * replaces is the body of the following reference functions (paths relative to the reference root):
 *
 *   boxLCD/world_env.py:431-458  WorldEnv.step       -> blcd_step / blcd_step_host / blcd_rollout
 *   boxLCD/world_env.py:446-450  b2World.Step x3     -> (inside blcd_step; Box2D 2.3.x, pybox2d 2.3.10)
 *   boxLCD/world_env.py:387-429  WorldEnv._get_obs   -> blcd_observe
 *   boxLCD/world_env.py:460-512  WorldEnv.lcd_render -> blcd_observe (frames), blcd_render_poses
 *   boxLCD/world_env.py:306-385  WorldEnv.reset      -> blcd_reset
 *   examples/collect.py:31-39    rollout loop        -> blcd_rollout
 *
 * Conventions
 *   - plain C, no torch types.  Pointers named *_dev are CUDA device pointers on the handle's device;
 *     pointers named *_host are ordinary host memory.  `stream` is a cudaStream_t passed as uint64
 *     (0 = legacy default stream).
 *   - every call returns 0 on success, <0 on error; blcd_last_error() returns a message for the
 *     calling thread's last failing call.
 *   - a handle is bound to one device and is not thread-safe; distinct handles may be driven from
 *     distinct threads / processes (one process per GPU is the intended multi-GPU layout).
 *   - LCD frames are bit-packed: one uint32 per output row, bit x = pixel x (1 = background,
 *     0 = body pixel, i.e. the reference's bool value), row 0 = top of the world (world_env.py:506).
 *     Frames wider than 32 px (lcd_base=32 scenes, envs.py:116-137: 64 x 32) take BLCD_LCD_WORDS(lcd_w) = 2
 *     words per row, word k holding pixels 32k .. 32k+31: every "[.., lcd_h] uint32" below reads
 *     "[.., lcd_h, 2] uint32" for them.
 *   - scenes are compiled for one of two size profiles, chosen by blcd_create from the spec: the
 *     small profile (<= 8 bodies, <= 7 joints, <= 64 collidable fixture pairs, frames <= 32 px wide:
 *     every env of envs.py:17-110) and the large profile (up to the BLCD_MAX_* limits, <= 128 pairs,
 *     frames <= 64 px wide: Crab / CrabCube / SpiderCube).  Same sources, same results; the limits
 *     only size registers, shared memory and per-world state.  BLCD_PROFILE=large forces the large one.
 *   - full_state / proprio are the reference's normalized observations (world_env.py:387-428,
 *     boxLCD/utils.py:119) as float32.
 */
#ifndef BOXLCD_B200_H
#define BOXLCD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLCD_MAX_BODIES 18  /* dynamic bodies per world (robot root + children + objects); CrabCube has 18 */
#define BLCD_MAX_JOINTS 17
#define BLCD_MAX_WALLS 4
#define BLCD_MAX_VERTS 8    /* b2_maxPolygonVertices */
#define BLCD_MAX_OBS (4 * BLCD_MAX_BODIES)
#define BLCD_LCD_WORDS(lcd_w) (((lcd_w) + 31) / 32)

enum { BLCD_SHAPE_CIRCLE = 0, BLCD_SHAPE_BOX = 1, BLCD_SHAPE_POLYGON = 2 };
enum { BLCD_ROLE_OBJECT = 0, BLCD_ROLE_ROOT = 1, BLCD_ROLE_CHILD = 2 };
enum { BLCD_RASTER_PIL12 = 0, BLCD_RASTER_PIL9 = 1 };
/* spec.flags: Box2D revision switches (0 = the 2.3.1+ forms; the host layer's default is BLCD_FLAG_DAMPING_2_3_0, which is
 * what pybox2d 2.3.10's recorded episodes show) and ablation switches */
enum {
  BLCD_FLAG_DAMPING_2_3_0 = 1,   /* v *= clamp(1 - h*d, 0, 1) instead of the Pade form 1/(1+h*d) */
  BLCD_FLAG_REFFACE_2_3_0 = 2,   /* b2CollidePolygons as in 2.3.0: hill-climbing b2FindMaxSeparation, reference-face rule
                                    0.98*sepA+0.001 (2.3.1+: brute-force search, sepA+0.1*linearSlop) */
  BLCD_FLAG_NO_TOI = 4,          /* continuousPhysics off (debug / ablation) */
  BLCD_FLAG_NO_SLEEP = 8         /* allowSleep off (debug / ablation) */
};

/* One fixture shape as the reference hands it to pybox2d (world_defs.py, world_env.py:272). Doubles are
 * the Python-side values; the library rounds to float32 exactly where Box2D would. */
typedef struct {
  int32_t kind;                       /* BLCD_SHAPE_* */
  int32_t n_verts;                    /* POLYGON: number of input vertices (hull is computed like b2PolygonShape::Set) */
  double radius;                      /* CIRCLE */
  double verts[BLCD_MAX_VERTS][2];    /* BOX: verts[0] = (hx, hy); POLYGON: input vertices */
} blcd_shape_def;

typedef struct {
  int32_t n_variants;                 /* 1; 2 for Object(shape='random'): variant 0 = circle, 1 = box (world_env.py:273-274) */
  int32_t role;                       /* BLCD_ROLE_* */
  blcd_shape_def shape[2];
  double density, friction, restitution, linear_damping, angular_damping;
  uint32_t category_bits, mask_bits;
  int32_t rand_angle;                 /* root: Robot.rand_angle; object: Object.rand_angle */
  int32_t parent;                     /* child: dynamic-body index of the joint's parent */
  int32_t root;                       /* child: dynamic-body index of its robot's root */
  double extent;                      /* root: Robot.bound; object: Object.size */
  double joint_angle;                 /* child: Joint.angle (relative to the ROOT angle, world_env.py:235) */
  double anchor_a[2], anchor_b[2];    /* child: Joint.anchorA / anchorB */
  int32_t obs_index[4];               /* index of this body's x:p, y:p, cos, sin in full_state (sorted keys) */
} blcd_body_def;

typedef struct {
  int32_t body_a, body_b;             /* dynamic-body indices (parent, child) */
  int32_t enable_limit, enable_motor;
  double anchor_a[2], anchor_b[2];
  double lower, upper;
  double max_motor_torque;            /* Joint.torque */
  double speed;                       /* Joint.speed: motorSpeed = speed * clip(action, -1, 1) (world_env.py:441) */
  int32_t act_index;                  /* index into the action vector, -1 if not actuated */
  int32_t _pad;
} blcd_joint_def;

typedef struct {
  int32_t n_bodies, n_joints, n_walls, has_robot;
  blcd_body_def bodies[BLCD_MAX_BODIES];   /* creation order = draw order: per robot root, children; then objects */
  blcd_joint_def joints[BLCD_MAX_JOINTS];  /* creation order */
  double walls[BLCD_MAX_WALLS][4];         /* static edge shapes x1,y1,x2,y2 (world_env.py:311-314), creation order */
  double gravity[2];
  int32_t world_w, world_h;                /* WIDTH = int(wh_ratio*base_dim), HEIGHT = base_dim (world_env.py:144-150) */
  int32_t lcd_w, lcd_h;                    /* frame size in pixels, lcd_w <= 64 (32 in the small profile, 64 in the large one) */
  int32_t obs_size, pobs_size, act_size;   /* pobs_size = 0 -> proprio is zeros(1) */
  int32_t pobs_index[BLCD_MAX_OBS];
  int32_t n_substeps;                      /* 3 when fps < 30 else 1 (world_env.py:446-452) */
  int32_t vel_iters, pos_iters;            /* 180, 60 */
  double dt;                               /* 1/(fps*3) or 1/fps */
  int32_t ep_len;
  int32_t raster_rules;                    /* BLCD_RASTER_* */
  uint32_t flags;                          /* BLCD_FLAG_* */
  int32_t _pad;
} blcd_spec;

typedef struct blcd_env* blcd_handle;

/* Number of floats per body in the raw body-state exchange format used by blcd_set_bodies/blcd_get_bodies:
 * x, y (body origin, i.e. b2Body.position), angle, vx, vy (of the centre of mass), omega. */
#define BLCD_BODY_STATE 6

const char* blcd_last_error(void);
int blcd_version(void);

/* Build the device-side tables for `spec` and allocate state for n_worlds worlds on `device`.
 * World w of this handle has global index world_offset + w: RNG streams are keyed by (seed, global index),
 * so results do not depend on how worlds are sharded over handles / GPUs. */
int blcd_create(const blcd_spec* spec_host, int64_t n_worlds, int device, uint64_t seed, int64_t world_offset, blcd_handle* out);
int blcd_destroy(blcd_handle h);

/* Re-key the random streams of an existing handle: world w becomes global world world_offset + w of stream family `seed`.
 * Nothing is reallocated; the next blcd_reset samples from the new streams.  This is what lets one allocation walk over a
 * long dataset (the reference's collector re-uses its worker processes across barrels, research/data.py:36-79). */
int blcd_rekey(blcd_handle h, uint64_t seed, int64_t world_offset);

/* reset(): idx_dev = NULL resets all worlds (n ignored), else the n listed worlds.  full_state_dev = NULL samples the
 * reference's reset distribution (world_env.py:197-304) from the per-world RNG; otherwise [n, obs_size] normalized
 * states are applied on top of a sampled reset exactly as reset(full_state=...) does (world_env.py:323-380). */
int blcd_reset(blcd_handle h, const int64_t* idx_dev, int64_t n, const float* full_state_dev, uint64_t stream);

/* One env.step() for every world: motor speeds from actions (clip to [-1,1] * Joint.speed), n_substeps x b2World.Step,
 * ep_t += 1.  actions_dev = NULL draws a ~ U[-1,1) float32 per action dim from the per-world RNG (collect.py:35);
 * actions_out_dev (optional, [N, act_size]) receives the actions used. */
int blcd_step(blcd_handle h, const float* actions_dev, float* actions_out_dev, uint64_t stream);

/* _get_obs() for every world.  Any output pointer may be NULL.  full_state [N, obs_size]; proprio [N, max(pobs_size,1)];
 * lcd_bits [N, lcd_h] uint32; lcd_bool [N, lcd_h, lcd_w] uint8 (reference layout); done [N] uint8 (ep_t >= ep_len). */
int blcd_observe(blcd_handle h, float* full_state_dev, float* proprio_dev, uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev,
                 uint8_t* done_dev, uint64_t stream);

/* blcd_step immediately followed by blcd_observe in the same kernel (one launch, state stays on chip in between);
 * this is what WorldEnv.step returns (world_env.py:458).  Any output pointer may be NULL. */
int blcd_step_observe(blcd_handle h, const float* actions_dev, float* actions_out_dev, float* full_state_dev, float* proprio_dev,
                      uint32_t* lcd_bits_dev, uint8_t* lcd_bool_dev, uint8_t* done_dev, uint64_t stream);

/* collect.py's inner loop, device resident: for t in [0,T): record obs_t (full_state, lcd bits), draw a_t on device,
 * record it, step.  Outputs are [N, T, ...] (world-major, the npz layout).  Worlds must have been reset by the caller. */
int blcd_rollout(blcd_handle h, int32_t T, float* full_state_dev, uint32_t* lcd_bits_dev, float* actions_dev, uint64_t stream);

/* Host-buffer variant of step + observe (the call an unmodified reference-side caller would bind): copies actions
 * host->device, steps, observes, copies full_state and packed frames device->host, and synchronizes.
 * Buffers that lie inside memory the caller page-locked with blcd_pin_host are copied to / from directly (DMA);
 * any other buffer is staged through pinned memory owned by the handle, so ordinary pageable arrays (a fresh numpy
 * array per step) are safe to pass.  The call never page-locks caller memory on its own. */
int blcd_step_host(blcd_handle h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host);

/* The same split into submit and wait, the call shape of the reference's AsyncVectorEnv.step_async / step_wait
 * (research/wrappers/async_vector_env.py:131-242), for callers whose next actions do not depend on the step just
 * submitted (collect.py:35 samples actions independently of the observation; double-buffered RL):
 *   blcd_pin_host         page-locks caller memory once (a whole [T, N, ...] dataset array, or a ring of step buffers);
 *                         the memory must stay allocated until blcd_unpin_host(same address) or blcd_destroy;
 *   blcd_step_host_async  enqueues one env step reading / writing buffers inside pinned memory and returns at once
 *                         (up to 3 further steps may be queued behind it; it blocks only when that queue is full);
 *   blcd_step_host_wait   returns when at most keep_in_flight submitted steps are unfinished; the outputs of every
 *                         finished step are then complete in host memory.  Steps finish in submission order.
 * Each step still copies its actions host->device and its observations device->host; consecutive steps overlap only in
 * that the copies and the ragged kernel tail of one step hide behind the next step's kernel. */
int blcd_pin_host(blcd_handle h, const void* buf_host, int64_t bytes);
int blcd_unpin_host(blcd_handle h, const void* buf_host);
int blcd_step_host_async(blcd_handle h, const float* actions_host, float* full_state_host, uint32_t* lcd_bits_host, uint8_t* done_host);
int blcd_step_host_wait(blcd_handle h, int32_t keep_in_flight);

/* lcd_render() from explicit poses, no simulation state involved: poses_dev [N, n_bodies, 4] = (x, y, sin, cos) float32
 * of every dynamic body's b2Transform; variant_dev optional [N] uint32 bitmask selecting shape variant per body. */
int blcd_render_poses(blcd_handle h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, uint32_t* lcd_bits_dev, uint64_t stream);

/* Same at an explicit frame size, lcd_render(width, height) (world_env.py:460-470); 0 = the scene's own size.  Any width up to
 * 1024 (frames wider than the profile's row mask are rendered as 32- or 64-pixel column windows, one launch each);
 * output rows have BLCD_LCD_WORDS(lcd_w) words; lcd_h <= 256.  The reference's viewer asks for 8x the env's size. */
int blcd_render_poses_sized(blcd_handle h, const float* poses_dev, const uint32_t* variant_dev, int64_t n, int32_t lcd_w, int32_t lcd_h,
                            uint32_t* lcd_bits_dev, uint64_t stream);

/* Raw body state, [N, n_bodies, BLCD_BODY_STATE] float32.  blcd_set_bodies puts every world into the state of a freshly
 * built b2World with bodies at those poses/velocities (no contacts yet, zero warm-start impulses, ep_t = 0), which is
 * the protocol single-step parity is defined on (last paragraph). */
int blcd_set_bodies(blcd_handle h, const float* bodies_dev, const uint32_t* variant_dev, uint64_t stream);
int blcd_get_bodies(blcd_handle h, float* bodies_dev, uint64_t stream);
/* b2Transform of every dynamic body as lcd_render consumes it: poses [N, n_bodies, 4] = (x, y, sin, cos); variant_dev
 * (optional, [N]) receives the shape-variant bitmask.  Feeding both to blcd_render_poses reproduces blcd_observe's frames. */
int blcd_get_poses(blcd_handle h, float* poses_dev, uint32_t* variant_dev, uint64_t stream);

/* Checkpoint / restore of the complete per-world simulation state (poses, velocities, warm-start impulses, contact
 * slots, RNG counters).  blcd_state_bytes() gives the buffer size for this handle. */
int64_t blcd_state_bytes(blcd_handle h);
int blcd_save_state(blcd_handle h, void* buf_dev, uint64_t stream);
int blcd_load_state(blcd_handle h, const void* buf_dev, uint64_t stream);

/* Failure detection: counts the worlds whose body state is no longer finite (NaN / inf after a diverged solve) and, if
 * invalid_dev != NULL ([N] uint8), flags them; the caller decides whether to blcd_reset those indices.  Synchronous. */
int blcd_check_finite(blcd_handle h, uint8_t* invalid_dev, int64_t* n_invalid_host);

/* Measured denominators for the solver roofline (bench.py): out4 = { lane-FMAs per second of the whole device with every
 * issue slot filled (x2 = fp32 FLOP/s; = peak lane-instructions per second), FFMAs per second of ONE dependent chain
 * (clock / dependent-issue latency: the regime of a serial Gauss-Seidel sweep), SM count, nominal max SM clock in Hz }.
 * Synchronous; takes ~0.1 s. */
int blcd_measure_peaks(int device, double* out4);

/* Introspection used by tests and bench.py */
int64_t blcd_num_worlds(blcd_handle h);
int64_t blcd_kernel_launches(blcd_handle h);   /* number of kernels this handle has launched so far */
int blcd_last_step_ms(blcd_handle h, float* ms_out); /* CUDA-event duration of the most recent step kernel (timing must have been enabled) */
int blcd_enable_timing(blcd_handle h, int on);
/* per-world diagnostic counters accumulated by blcd_step: [N, BLCD_N_COUNTERS] uint32 */
#define BLCD_N_COUNTERS 8
enum { BLCD_CNT_CONTACTS = 0, BLCD_CNT_POS_ITERS = 1, BLCD_CNT_TOI_EVENTS = 2, BLCD_CNT_TOI_CALLS = 3, BLCD_CNT_SLEEP_STEPS = 4,
       BLCD_CNT_OVERFLOW = 5, BLCD_CNT_MANIFOLD_POINTS = 6, BLCD_CNT_SUBSTEPS = 7 };
int blcd_get_counters(blcd_handle h, uint32_t* counters_dev, uint64_t stream);
/* out16: n_bodies, n_joints, n_walls, n_pairs, obs_size, pobs_size, act_size, lcd_w, lcd_h, manifold slots per world,
 * state words per world, shared-memory words per world, threads per block (fused kernels), shared-memory bytes per block,
 * scene-size profile (0 small, 1 large), device path (0 fused one-thread-per-world kernel, 1 phase pipeline) */
int blcd_scene_info(blcd_handle h, int32_t* out16);

#ifdef __cplusplus
}
#endif
#endif /* BOXLCD_B200_H */
