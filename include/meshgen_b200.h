/*
 * meshgen_b200 -- C ABI of the B200-native batched BoudaryEnv (quad-mesh generation RL env).
 *
 * This is the drop-in boundary of the hot path.  The reference has no
 * native interface: its boundary is the Python class BoudaryEnv.  Each entry point below names
 * the reference method it replaces; citations are relative to the reference tree,
 *   E = v2/src/mesh_rl/envs/boundary_env.py   (legacy twin: rl/boundary_env.py)
 *   M = v2/src/mesh_rl/mesh_core.py           (legacy twin: general/mesh.py)
 *   C = v2/src/mesh_rl/components_core.py     (legacy twin: general/components.py)
 *   V = rl/baselines/dummy_vec_env.py         (the VecEnv surface the trainers use)
 *
 * Conventions
 *   - plain C types only; every call returns 0 on success, a negative mg_status otherwise, and
 *     never throws.  mg_last_error() gives a human-readable message for the last failure.
 *   - a handle owns all device state of `num_envs` environments on one CUDA device.
 *   - pointers named *_dev are device pointers owned by the caller (e.g. tensor.data_ptr());
 *     pointers named *_host are host pointers.  `stream` is a cudaStream_t passed as void*.
 *   - calls enqueue work on `stream` and do not synchronise unless stated.
 *   - one handle is not thread-safe; different handles are independent.
 *   - there is no CPU fallback: without a CUDA device mg_create fails with MG_ERR_CUDA.
 */
#ifndef MESHGEN_B200_H
#define MESHGEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_OBS_DIM 18 /* E:91-94  2 * (neighbor_num 6 + radius_num 3) float32 */
#define MG_ACT_DIM 3  /* E:78-80  Box([-1,-1.5,0],[1,1.5,1.5]) float32 */

typedef struct mg_env_s *mg_handle;

typedef enum mg_status {
    MG_OK = 0,
    MG_ERR_ARG = -1,    /* bad argument */
    MG_ERR_CUDA = -2,   /* CUDA runtime failure (message has the cudaError string) */
    MG_ERR_STATE = -3,  /* call order (e.g. step before domains were set) */
    MG_ERR_CAPACITY = -4 /* polygon larger than max_verts */
} mg_status;

/* Random star-polygon generator configuration (ui/GenerateRandomPolygon.py:5-49, defaults :63;
 * densifier ui/tk-ui.py:252-276).  Lengths in pixel units of the reference tools; the
 * generated coordinates are divided by 100 like read_polygon does (geometry.py:46). */
typedef struct mg_polygen_cfg {
    double ctr_x, ctr_y;     /* 250, 250 */
    double ave_radius;       /* 100 */
    double irregularity;     /* 0.55 */
    double spikeyness;       /* 0.7 */
    int32_t min_coarse;      /* 8   coarse vertices K ~ U{min_coarse..max_coarse} */
    int32_t max_coarse;      /* 24 */
    int32_t min_verts;       /* 64  densified vertex count n, even, min_verts <= n <= max_verts */
    int32_t max_verts;       /* 512 */
} mg_polygen_cfg;

/* Episode statistics summed over all envs of the handle since the last mg_stats(reset=1).
 * This 12-element vector (10 int64 counters, 2 float64 sums) is the only thing the multi-GPU driver all-reduces. */
typedef struct mg_episode_stats {
    int64_t episodes;      /* finished episodes */
    int64_t completed;     /* ... that ended with is_complete (terminated) */
    int64_t truncated;     /* ... that ended by 100 consecutive failures (E:382-384) */
    int64_t steps;         /* env steps executed */
    int64_t successes;     /* steps that created an element */
    int64_t elements;      /* elements in finished episodes */
    int64_t sum_n;         /* sum over steps of the live boundary size (roofline accounting) */
    int64_t sum_n_success; /* same, restricted to successful steps */
    int64_t ring_items;    /* steps that needed the whole boundary (work items of the ring kernel) */
    int64_t sum_n_ring;    /* sum of the live boundary size over those steps */
    double sum_return;     /* sum of finished-episode returns */
    double sum_length;     /* sum of finished-episode lengths */
} mg_episode_stats;

/* Read-back of one env's state for parity tests (host buffers sized by max_verts). */
typedef struct mg_state_view {
    int32_t n;              /* live boundary size, len(updated_boundary.vertices) */
    int32_t ref_index;      /* index of the reference point in the boundary list, -1 if none */
    int32_t n_elements;     /* len(generated_meshes) */
    int32_t failed_num;     /* E:378-384 */
    int32_t n0;             /* original polygon size */
    int32_t memo_flags;     /* memoised verdict of the rule -1 / +1 elements in this state: bit 0 / 1 accepted, bit 2 / 3
                             * valid with the boundary-intersection test still pending (see mg_step) */
    double base_length;     /* C:1089-1090 */
    double current_area;    /* E:336 */
    double original_area;   /* E:72 */
    double area_min, area_crit; /* M:705-718 estimated_area_range */
    double *xy_host;        /* out, 2*max_verts doubles (x0,y0,x1,y1,...) or NULL */
    int32_t *vertex_id_host;/* out, max_verts: 0..n0-1 original, n0+k k-th inserted vertex */
    double *cand_key_host;  /* out, max_verts: candidate key (degrees) or +inf when not a candidate */
    int32_t *cand_stamp_host;/* out, max_verts: tie-break stamp (smaller = earlier in the list) */
} mg_state_view;

/* Replaces BoudaryEnv.__init__ (E:58-120) for a batch: allocates state for num_envs
 * environments whose polygons have at most max_verts vertices, on CUDA device `device`. */
int mg_create(mg_handle *out, int device, int num_envs, int max_verts);

/* Replaces BoudaryEnv(boundary) / from_domain_file (E:45-56) + read_polygon (geometry.py:34-52):
 * n_domains polygons given clockwise as packed (x, y) float64 pairs; polygon d spans
 * xy_host[2*offsets_host[d] .. 2*offsets_host[d+1]).  env_domain_host[e] selects the polygon of
 * env e.  areas_host (optional, may be NULL) carries Boundary2D.poly_area() (C:485-487) per
 * domain as computed by the caller; when NULL the library evaluates the shoelace sum itself.
 * Computes the per-domain reset template on the device (candidate keys M:228-287, area range
 * M:705-718, first observation C:1192-1290) and synchronises. */
int mg_set_domains(mg_handle h, const double *xy_host, const int32_t *offsets_host, int n_domains,
                   const int32_t *env_domain_host, const double *areas_host);

/* Workload generator for BASELINE configs 3/4: every reset draws a fresh random star polygon
 * in-kernel (counter-based Philox stream: seed, subsequence = global env id, so results do not
 * depend on how envs are sharded over ranks).  env_id_offset = first global env id of this handle. */
int mg_set_random(mg_handle h, uint64_t seed, const mg_polygen_cfg *cfg, int64_t env_id_offset);

/* enabled != 0 (default): VecEnv convention, a finished env is reset in place by mg_step and
 * the pre-reset observation goes to term_obs.  enabled == 0: plain Gym env semantics
 * (E:388-457): the env stays in its final state until mg_reset; obs_dev holds the final
 * observation, stepping a finished env repeats the reference's behaviour (reward 10, done). */
int mg_set_auto_reset(mg_handle h, int enabled);

/* Replaces BoudaryEnv.reset (E:136-184).  mask_dev: num_envs bytes (non-zero = reset that env)
 * or NULL for all.  obs_dev: out, num_envs * 18 float32 (all envs' current observation). */
int mg_reset(mg_handle h, const uint8_t *mask_dev, float *obs_dev, void *stream);

/* Replaces BoudaryEnv.step (E:388-457) followed by the VecEnv auto-reset (V:40-52 with the
 * stock SB3 behaviour): one transition for every env.  Four launches: a screen kernel (one thread per env) that
 * settles every step whose outcome follows from the env's 128-byte record -- the verdict of the rule -1 / +1
 * elements depends on the state only and is memoised, and a new-vertex element that fails Mesh.is_valid
 * (C:738-757) fails whatever the point-in-polygon test says -- and three warp-per-item kernels (decide, update,
 * observe) for the steps that need the whole boundary.  Results are the reference's for every step.
 *   act_dev      in  num_envs*3 float32
 *   obs_dev      out num_envs*18 float32   next observation (after auto-reset when done)
 *   rew_dev      out num_envs float64      reward (the reference returns np.float64)
 *   term_dev     out num_envs uint8        terminated = done and is_complete
 *   trunc_dev    out num_envs uint8        truncated  = done and not is_complete
 *   term_obs_dev out num_envs*18 float32   observation before the auto-reset; rows are written only
 *                                          where done (others keep their content); may be NULL
 *   n_elem_dev   out num_envs int32        len(generated_meshes) after the step, before reset
 *                                          (may be NULL) */
int mg_step(mg_handle h, const float *act_dev, float *obs_dev, double *rew_dev, uint8_t *term_dev,
            uint8_t *trunc_dev, float *term_obs_dev, int32_t *n_elem_dev, void *stream);

/* Replaces BoudaryEnv.move (E:459-594; legacy rl/boundary_env.py:265-432): the "apply this geometric move" entry point of
 * the reference's data-generation utilities (general/EBRD.py, general/FNN_evaluation.py), one move for every env.
 *   polar_dev     in  num_envs*2 float64  new_point = (r, phi), in units of radius (4) x base_length, Python floats in the
 *                                         reference (rounded to 6 decimals by CPython's round, E:478)
 *   type_dev      in  num_envs float64    <= 0.3: rule -1 element, >= 0.7: rule +1 element, otherwise a new vertex
 *   obs_dev       out num_envs*18 float32 next observation of a STATIC point environment (area-ratio slot 0, C:1209-1214);
 *                                         zeros where the reference returns None
 *   done_dev / complete_dev out uint8     done and info["is_complete"] as the reference returns them (reward is always 0)
 *   exhausted_dev out uint8               every reference candidate was on the not-valid list and the mesh could NOT be
 *                                         smoothed (see below): the env reports done.  Reset it before the next move.
 *   n_elem_dev    out num_envs int32      len(generated_meshes) (may be NULL)
 * A failed move puts the reference point on the env's not-valid list (cleared by the next accepted element or by
 * mg_reset), and the next reference point is the first candidate not within 0.001 of a listed point (M:310-314,
 * M:428-433).  mg_step does not look at that list (in the reference, step() after failed move()s would): reset between
 * the two APIs.  Entering with n <= 5 raises in the reference (unbound is_complete); here: done, complete iff n <= 4.
 * When every candidate is on the list the reference smooths the whole mesh (smooth_pave, general/mesh.py:790-1067,
 * 1258-1288) and goes on (E:548-583): so does mg_move, on the envs the move kernel reports -- it reads their
 * flags, i.e. synchronises `stream` once per call.  Smoothed coordinates agree with the reference to <= 1e-9 (device libm
 * vs glibc), every discrete outcome exactly.  mg_set_option("smooth_pave", 0) turns it off. */
int mg_move(mg_handle h, const double *polar_dev, const double *type_dev, float *obs_dev, uint8_t *done_dev, uint8_t *complete_dev,
            uint8_t *exhausted_dev, int32_t *n_elem_dev, void *stream);

/* Same transition through HOST buffers (what a numpy-facing caller such as SB3's VecEnv pays):
 * H2D of the actions, the step, D2H of all outputs, one stream synchronise.  Buffers that are pinned
 * (cudaHostAlloc / cudaHostRegister / torch .pin_memory()) are used by the step kernels directly -- the actions are
 * read over PCIe by the first kernel while it loads the env records, the outputs are written by the kernels that
 * produce them; pageable ones go through staging buffers and one copy each.  term_obs_host rows are defined only
 * where the episode ended.
 * Runs on a private stream after everything the caller enqueued through mg_reset / mg_step / mg_snapshot_* has
 * finished, and returns after its own work has finished. */
int mg_step_host(mg_handle h, const float *act_host, float *obs_host, double *rew_host,
                 uint8_t *term_host, uint8_t *trunc_host, float *term_obs_host, int32_t *n_elem_host);
/* The two halves of mg_step_host: _begin enqueues the step on the handle's private stream and returns, _end waits for
 * it (the result arrays are defined after _end).  For callers that split their envs over two handles (SB3's VecEnv
 * step_async / step_wait, rl/baselines/dummy_vec_env.py:38-58, is the same split): while one half runs on the GPU the
 * host reads the other half's results, runs its policy and enqueues its next step, so the host's latency and the PCIe
 * transfers of one half hide behind the other half's kernels.  One step per handle in flight: until _end, a second
 * _begin and the calls that touch the env state (mg_reset, mg_step, mg_move, mg_snapshot_*) are refused with
 * MG_ERR_STATE. */
int mg_step_host_begin(mg_handle h, const float *act_host, float *obs_host, double *rew_host,
                       uint8_t *term_host, uint8_t *trunc_host, float *term_obs_host, int32_t *n_elem_host);
int mg_step_host_end(mg_handle h);

/* Observation delta.  A failed step leaves the env untouched and the reference itself returns a bit-identical
 * observation, so with enabled != 0 mg_step / mg_step_host write only the observation rows that changed (accepted
 * element or reset) when they are handed the SAME obs buffer as the previous call (or the preceding mg_reset) --
 * the caller promises not to modify that buffer in between.  A different pointer, mg_reset without an obs buffer or
 * mg_snapshot_load fall back to one full write.  Off by default.  mg_set_host_delta is the round-1 name.
 * mg_step_host extends the same promise to PINNED reward / flag / element-count arrays: a value crosses PCIe only where it
 * differs from what the caller's array already holds (a failed step mostly repeats the env's previous reward, flags and
 * count), so do not modify those arrays between calls either -- read them, or copy them. */
int mg_set_obs_delta(mg_handle h, int enabled);
int mg_set_host_delta(mg_handle h, int enabled);

/* Bytes moved host->device and device->host by the last mg_step_host call (evaluated here, from the step counters of
 * that call: synchronises). */
int mg_last_host_bytes(mg_handle h, int64_t *h2d, int64_t *d2h);

/* Uniform actions in the action box (E:78-80) from the handle's Philox stream -- the synthetic
 * policy used by the benchmarks.  act_dev: out num_envs*3 float32. */
int mg_sample_actions(mg_handle h, uint64_t seed, uint64_t step_index, float *act_dev, void *stream);
/* The same with the step index in device memory: step_counter_dev[0] = the index this launch uses, advanced by one when
 * the launch is done (step_counter_dev[1] is scratch, zero it once), so that policy + mg_step can be captured in a CUDA
 * graph once and replayed -- every replay draws the next step's actions. */
int mg_sample_actions_seq(mg_handle h, uint64_t seed, uint64_t *step_counter_dev, float *act_dev, void *stream);

/* Parity/debug read-back of env `env` (synchronises). */
int mg_get_state(mg_handle h, int env, mg_state_view *view);

/* Element log of env `env` (generated_meshes, E:331,351): up to max_elements quads as 4 vertex
 * ids each, and the coordinates of every vertex id < n_vertices (original polygon first; in random-polygon mode it
 * is regenerated from the episode's counter).  Synchronises.  Returns the number of elements through
 * *n_elements_out and of vertices through *n_vertices_out.  MG_ERR_CAPACITY when the episode outgrew the log (the
 * counts and the returned prefix are still valid). */
int mg_get_elements(mg_handle h, int env, int32_t *quads_host, int max_elements, int32_t *n_elements_out,
                    double *vertex_xy_host, int max_vertices, int32_t *n_vertices_out);

/* Parity / debug read-back of the workload generator (random-polygon mode; synchronises): the polygon of episode
 * `episode` of env `env` (episode < 0: the env's current episode) regenerated from its Philox counter.
 *   xy_host        out, up to max_vertices (x, y) pairs of the densified clockwise ring (env units); *n_out its size
 *   area_out       out, the shoelace area the in-kernel reset computes for it (original_area)
 *   coarse_px_host out, 64 int32: the coarse star polygon in pixels, clockwise (ui/GenerateRandomPolygon.py:5-49 reversed);
 *                  *k_out its vertex count;  *spacing_out the densifier spacing A in pixels (ui/tk-ui.py:252-276)
 * Any output pointer may be NULL. */
int mg_debug_polygon(mg_handle h, int env, int episode, double *xy_host, int max_vertices, int32_t *n_out, double *area_out,
                     int32_t *coarse_px_host, int32_t *k_out, double *spacing_out);

/* Capacity of the per-env element log and inserted-vertex log (what mg_get_elements can return; the element COUNT is
 * always exact).  Default: 2 x max_verts each (at least 64).  Element counts scale with the domain's area: the
 * reference's evaluation runs report up to ~5 x n0 elements per episode (rl/baselines/evaluation.txt), so an
 * evaluator raises it (the Python BoudaryEnv facade and evaluation loop use 8 x).  mg_set_log_capacity reallocates
 * the logs (call it before mg_reset; contents are discarded). */
int mg_set_log_capacity(mg_handle h, int max_elements_per_env, int max_inserted_per_env);
int mg_log_capacity(mg_handle h, int32_t *max_elements_per_env, int32_t *max_inserted_per_env);

/* Sum the per-env episode counters on the device, copy them to *out (synchronises);
 * reset != 0 zeroes the counters afterwards. */
int mg_stats(mg_handle h, mg_episode_stats *out, int reset);

/* Same sum written to a DEVICE mg_episode_stats on `stream`, no synchronisation: what a multi-GPU job all-reduces
 * every few steps (10 int64 counters followed by 2 float64 sums) without stalling the step stream. */
int mg_stats_async(mg_handle h, mg_episode_stats *stats_dev, int reset, void *stream);

/* Device-resident replay buffer write (replaces, for a batched env, what SB3's
 * OffPolicyAlgorithm._store_transition + ReplayBuffer.add do on the host -- the callers named in
 * rl/baselines/RL_Mesh.py:186-197 and v2 training/train_loop.py:132): one launch stores the N transitions of a
 * step into slot `slot` of a ring of `capacity_steps` slots, layout [slot][env][...]:
 *   buf_obs[slot][e]      = prev_obs[e]              (observation the action was computed from)
 *   buf_next_obs[slot][e] = done ? term_obs[e] : new_obs[e]   (SB3 stores the terminal observation, not the reset one)
 *   buf_act[slot][e]      = act[e];  buf_rew[slot][e] = (float) rew[e]
 *   buf_done[slot][e]     = terminated | truncated;  buf_timeout[slot][e] = truncated
 *                           (handle_timeout_termination: the learner bootstraps unless done & !timeout)
 * All pointers are device pointers; work is enqueued on `stream`; no synchronisation. */
int mg_replay_add(mg_handle h, int64_t capacity_steps, int64_t slot, float *buf_obs, float *buf_next_obs, float *buf_act,
                  float *buf_rew, uint8_t *buf_done, uint8_t *buf_timeout, const float *prev_obs, const float *act,
                  const float *new_obs, const double *rew, const uint8_t *term, const uint8_t *trunc,
                  const float *term_obs, void *stream);

/* Snapshot / restore of the complete environment state (boundary rings, candidate keys, per-env scalars, element
 * log, statistics, cached observations) -- the reference never checkpoints its env; with a
 * batched env this is what makes long rollouts resumable and lets a caller branch from a state.  The blob is
 * mg_snapshot_bytes(h) bytes of DEVICE memory owned by the caller; copies are enqueued on `stream`.  A blob can be
 * loaded into any handle created with the same num_envs / max_verts and the same domains or generator settings; the header of the blob
 * records num_envs, max_verts, mode, log capacities, domain count, generator seed and env id offset, and
 * mg_snapshot_load rejects a blob that disagrees with the handle or is shorter than mg_snapshot_bytes.  The not-valid
 * lists of mg_move are not part of a snapshot (take it at an episode boundary or after an accepted move). */
int64_t mg_snapshot_bytes(mg_handle h);
int mg_snapshot_save(mg_handle h, void *blob_dev, void *stream);
int mg_snapshot_load(mg_handle h, const void *blob_dev, int64_t blob_bytes, void *stream);

/* Switches.  "smooth_pave" (default 1) is behaviour: 0 = mg_move stops where the reference would smooth.  The others are
 * tuning switches (results never depend on them).  "fuse_decide" (default 1): the decide and update work of a step
 * share one launch -- a warp that accepts an element applies it with the boundary it has already staged; 0 = two
 * launches with smaller code images.  "reset_side" (default 1): the in-place resets of the envs a step truncated run
 * in their own kernel on a side stream next to the update kernel (0: in the caller's stream, before the observe
 * kernel).  "pdl" (default 1): the update and observe kernels are launched as programmatic dependents of their
 * predecessor (their blocks are placed while it drains and wait for it with griddepcontrol.wait).  "update_blocks" /
 * "observe_blocks" / "reset_blocks": resident one-warp blocks per SM of those kernels (default: what fits). */
int mg_set_option(mg_handle h, const char *name, int value);

/* Profiling aid (bench.py roofline.per_kernel): with enabled != 0 every mg_step / mg_step_host records CUDA events
 * around its four kernels; mg_kernel_times synchronises, returns the mean device time in ms of the screen, decide,
 * update and observe kernels (ms4[0..3]; decide is 0 when it is fused into update) over the (at most 256 most recent) steps since the last call, and starts a
 * new window. */
int mg_set_kernel_timing(mg_handle h, int enabled);
int mg_kernel_times(mg_handle h, double *ms4, int64_t *steps);

int mg_num_envs(mg_handle h);
int mg_max_verts(mg_handle h);
/* number of kernel launches issued by this handle so far (bench.py's gpu_launches) */
int64_t mg_launch_count(mg_handle h);

int mg_destroy(mg_handle h);
const char *mg_last_error(mg_handle h);
const char *mg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MESHGEN_B200_H */
