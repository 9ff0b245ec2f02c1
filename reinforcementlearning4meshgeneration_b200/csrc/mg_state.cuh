// Device-resident state of the batched BoudaryEnv and the kernel parameter block.
// Layout in HBM (env-major, `cap` = max_verts rounded up to a multiple of 2):
//   xy    double2[num_envs][cap]   boundary vertex ring (updated_boundary.vertices), index 0 first
//   key   double [num_envs][cap]   cached candidate key in degrees (M:228-257), +inf = not a candidate
//   stamp int32  [num_envs][cap]   tie-break stamp: list index at rebuild, decreasing negatives later
//   vid   int32  [num_envs][cap]   vertex ids (0..n0-1 original, n0+k k-th inserted)
//   st    EnvState[num_envs]       scalars (128 B)
//   stats StatsAcc[64]             episode counters (fire-and-forget atomics, 64 slots)
//   obs   float  [num_envs][18]    cached observation (failed steps return it unchanged)
#pragma once
#include <stdint.h>

#include "../../include/meshgen_b200.h"

namespace mg {

// The first 48 bytes are everything phase A (the all-envs kernel) reads, and the second and third
// 16-byte chunks are everything it writes back on a failed step.
struct __align__(16) EnvHot {
    int32_t n;            // live boundary size
    int32_t ref_index;    // index of the reference point, -1 = none
    int32_t n_elements;   // len(generated_meshes)
    int32_t n0;           // size of the episode's original polygon
    double base_length;
    int32_t failed_num;   // consecutive failed steps
    int32_t ep_len;
    double ep_return;
    double current_area;
};
static_assert(sizeof(EnvHot) == 48, "EnvHot size");

struct __align__(16) EnvState {
    int32_t n;
    int32_t ref_index;
    int32_t n_elements;
    int32_t n0;
    double base_length;
    int32_t failed_num;
    int32_t ep_len;
    double ep_return;
    double current_area;
    // ---- not touched by phase A ----
    double original_area;
    double area_min;
    double area_crit;
    int32_t next_vid;     // id of the next inserted vertex
    int32_t stamp_ctr;    // decreasing stamp counter for incremental candidate inserts
    int32_t domain;       // template index (domain mode)
    int32_t episode;      // episodes finished by this env (random mode: polygon counter)
    int64_t pad[5];
};
static_assert(sizeof(EnvState) == 128, "EnvState size");

// Episode statistics: 64 accumulator slots updated with fire-and-forget atomics (slot = warp id & 63),
// summed on demand by mg_stats.  Same fields as mg_episode_stats.
constexpr int STAT_SLOTS = 64;
struct __align__(16) StatsAcc {
    unsigned long long episodes, completed, truncated, steps, successes, elements, sum_n, sum_n_success;
    double sum_return, sum_length;
    unsigned long long pad[6];       // 128 bytes: one slot per L2 line
};
static_assert(sizeof(StatsAcc) == 128, "StatsAcc size");

// element accepted by phase A of a step, applied by phase B
struct __align__(16) Pending {
    double newx, newy;    // the candidate vertex (used when new_vertex != 0)
    int32_t rule;         // -1 / +1 (0 with new_vertex)
    int32_t new_vertex;
    int64_t pad;
};

// counters[2 * set + {0, 1}] = sizes of the success / reset lists of counter set `set`; CNT_STEP = steps completed
// (set of a step = CNT_STEP & 1 when its phase A starts); CNT_CUR = the set phase A of the current step used
enum { CNT_STEP = 4, CNT_CUR = 5, CNT_N = 8 };

constexpr int ANGLE_TAB_N = 62833;      // round(2 pi, 4) = 6.2832

struct DomainScalars {
    int32_t n0;
    int32_t ref_index;
    double base_length;
    double original_area;
    double area_min;
    double area_crit;
};

struct Params {
    int num_envs;
    int cap;
    int auto_reset;      // 1 = VecEnv convention (reset in place when done), 0 = plain Gym env
    // env state
    double2 *xy;
    double *key;
    int32_t *stamp;
    int32_t *vid;
    EnvState *st;
    StatsAcc *stats;     // [STAT_SLOTS]
    float *obs_cache;
    // per-step work lists (phase kernels)
    Pending *pend;       // [num_envs]
    int *succ_list;      // [num_envs]
    int *reset_list;     // [num_envs]
    int *counters;       // [CNT_N]: two sets of list sizes + the device-side step parity (see the enum above)
    // element log: quads as 4 vertex ids, coordinates of inserted vertices
    int32_t *elem;       // [num_envs][elem_cap][4]
    double2 *ins_xy;     // [num_envs][ins_cap]
    int elem_cap;
    int ins_cap;
    // domain templates
    int n_domains;
    const double2 *t_xy;      // [n_domains][cap]
    const double *t_key;      // [n_domains][cap]
    const int32_t *t_stamp;   // [n_domains][cap]
    const DomainScalars *t_sc;
    const float *t_obs;       // [n_domains][18]
    // {sin, cos} of every quantised angle k * 1e-4 (k = 0..62832) and of its half, evaluated by the HOST libm at
    // mg_create: sin/cos on this path only ever see quantised angles (C:154-168, C:946-947, C:1243), and the
    // bisector ray test (C:657-676) is chaotic in the last bit of sin/cos on near-degenerate edges
    const double2 *sc_full;   // [ANGLE_TAB_N]
    const double2 *sc_half;   // [ANGLE_TAB_N]
    // random-polygon mode
    int random_mode;
    uint64_t seed;
    int64_t env_id_offset;
    mg_polygen_cfg gen;
};

}  // namespace mg
