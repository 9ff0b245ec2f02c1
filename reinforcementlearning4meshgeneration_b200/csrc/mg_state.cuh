// Device-resident state of the batched BoudaryEnv and the kernel parameter block.
// Layout in HBM (env-major, `cap` = max_verts rounded up to a multiple of 2):
//   xy    double2[num_envs][cap]   boundary vertex ring (updated_boundary.vertices), index 0 first
//   key   double [num_envs][cap]   cached candidate key in degrees (M:228-257), +inf = not a candidate
//   stamp int32  [num_envs][cap]   tie-break stamp: list index at rebuild, decreasing negatives later
//   vid   int32  [num_envs][cap]   vertex ids (0..n0-1 original, n0+k k-th inserted)
//   hot   EnvHot [num_envs]        128 B: everything the per-step screen kernel reads (scalars, the five
//                                  vertices around the reference point, the memoised rule -1 / +1 verdicts)
//   cold  EnvCold[num_envs]        64 B: per-episode constants and counters only the ring kernel needs
//   stats StatsAcc[64]             episode counters (fire-and-forget atomics, 64 slots)
//   obs   float  [num_envs][18]    cached observation (failed steps return it unchanged)
#pragma once
#include <stdint.h>

#include "../../include/meshgen_b200.h"

namespace mg {

// A step's outcome is a function of (state, action), and the state only changes when an element is accepted
// (~5 % of the steps under a random policy) or the env is reset.  Everything that depends on the state alone is
// therefore memoised in this record when the state changes:
//   * fan[]    the five vertices B[i-2..i+2] around the reference point i: the action frame (E:783-792), the
//              new-vertex quad [P, B[i-1], B[i], B[i+1]] and its Mesh.is_valid test (C:738-757) need nothing else
//   * flags    the rule -1 / rule +1 quad (E:236-283) is valid AND does not intersect the boundary
//              (C:738-757, M:536-556) -- both quads are made of boundary vertices only, so the verdict of a
//              rule -1 / +1 step does not depend on the action at all.  Mesh.is_valid (O(1)) is evaluated when the
//              state changes, the intersection scan (O(n)) the first time a rule action asks for it
// The screen kernel (one thread per env) settles every step whose outcome follows from this record; only the steps
// that need the whole boundary (point-in-polygon of a promising new vertex, accepted elements, resets) reach the
// ring kernel.
struct __align__(16) EnvHot {
    int32_t n;            // live boundary size
    int32_t ref_index;    // index of the reference point, -1 = none
    int32_t n_elements;   // len(generated_meshes)
    int32_t flags;        // HOT_OK_* | HOT_PEND_*
    double base_length;
    int32_t failed_num;   // consecutive failed steps
    int32_t ep_len;
    double ep_return;
    double current_area;
    double fan[10];       // B[i-2], B[i-1], B[i], B[i+1], B[i+2] as (x, y); valid when n >= 6 and ref_index >= 0
};
static_assert(sizeof(EnvHot) == 128, "EnvHot size");
// rule -1 / +1 element: accepted (OK), or valid with the boundary-intersection half of the verdict still to be
// evaluated (PEND); neither bit = the element is rejected in this state
enum { HOT_OK_M1 = 1, HOT_OK_P1 = 2, HOT_PEND_M1 = 4, HOT_PEND_P1 = 8 };

struct __align__(16) EnvCold {
    double original_area;
    double area_min;
    double area_crit;
    int32_t n0;           // size of the episode's original polygon
    int32_t next_vid;     // id of the next inserted vertex
    int32_t stamp_ctr;    // decreasing stamp counter for incremental candidate inserts
    int32_t domain;       // template index (domain mode)
    int32_t episode;      // episodes finished by this env (random mode: polygon counter)
    int32_t pad[5];
};
static_assert(sizeof(EnvCold) == 64, "EnvCold size");

// Episode statistics: 64 accumulator slots updated with fire-and-forget atomics, summed on demand by mg_stats.
// Same fields as mg_episode_stats.
constexpr int STAT_SLOTS = 64;
struct __align__(16) StatsAcc {
    unsigned long long episodes, completed, truncated, steps, successes, elements, sum_n, sum_n_success, ring_items, sum_n_ring;
    double sum_return, sum_length;
    unsigned long long pad[4];       // 128 bytes: one slot per L2 line
};
static_assert(sizeof(StatsAcc) == 128, "StatsAcc size");

// Work record handed from one step kernel to the next (32 bytes, so that the consumer needs no second dependent load
// before it can size the boundary copy).
//   WORK_DECIDE_NEW  : rule 0 with a promising candidate vertex (newx, newy): point-in-polygon etc. still open;
//                      flag = Mesh.is_valid of the new-vertex quad, already evaluated by the screen kernel
//   WORK_DECIDE_RULE : rule -1 / +1 whose quad is valid but whose boundary-intersection test is still pending
//   WORK_APPLY       : accepted element (rule, flag = new vertex at (newx, newy))
//   WORK_OBSERVE     : the env's state changed: next observation + memo (done = the episode completed)
//   WORK_RESET       : reset the env in place
struct __align__(16) WorkItem {
    double newx, newy;
    int32_t env;
    int32_t n;            // live boundary size (what to stage)
    int8_t kind;
    int8_t rule;          // -1 / +1 (0 = new vertex)
    int8_t flag;
    int8_t done;
    int32_t pad;
};
static_assert(sizeof(WorkItem) == 32, "WorkItem size");
enum { WORK_DECIDE_NEW = 0, WORK_DECIDE_RULE = 1, WORK_APPLY = 2, WORK_OBSERVE = 3, WORK_RESET = 4 };

// Every work list has NBINS segments of num_envs records per item kind, binned by boundary size (largest first), so that
// the item kernels hand out long items before short ones.  The decide list has two kinds (new-vertex candidates, then
// pending rule verdicts): the update kernel serves the lists kind by kind, so that the warps that share an SM -- and
// its instruction cache -- mostly run the same code at the same time.
constexpr int NBINS = 4;
constexpr int DECIDE_SEGS = 2 * NBINS;
// counters[CNT_SET * set + ...]: sizes of the decide / accept / observe list segments and of the reset list of counter
// set `set`, the item tickets of the warp-per-item kernels and the number of episodes that ended in the step; CNT_STEP =
// parity of the next step (the set its screen kernel will use); CNT_CUR = the set the current step uses.  Two sets
// alternate so that no memset sits between the launches of a step and any sequence of steps can be captured in a CUDA
// graph.
enum { CNT_DECIDE = 0, CNT_ACCEPT = 8, CNT_OBSERVE = 12, CNT_DONE = 16, CNT_TICKET_DECIDE = 17, CNT_TICKET_UPDATE = 18,
       CNT_TICKET_OBSERVE = 19, CNT_RESET = 20, CNT_TICKET_RESET = 21, CNT_RESBYTES = 22 /* result bytes written in delta mode */,
       CNT_SET = 24, CNT_STEP = 48, CNT_CUR = 49, CNT_N = 56 };

constexpr int ANGLE_TAB_N = 62833;      // round(2 pi, 4) = 6.2832

// Reset template of one domain (domain mode): the records of a freshly reset env.
struct DomainProto {
    EnvHot hot;
    EnvCold cold;
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
    EnvHot *hot;
    EnvCold *cold;
    StatsAcc *stats;     // [STAT_SLOTS]
    float *obs_cache;
    // per-step work lists
    WorkItem *decide_list;   // [DECIDE_SEGS][num_envs]  screen -> decide
    WorkItem *accept_list;   // [NBINS][num_envs]  screen / decide -> update
    WorkItem *observe_list;  // [NBINS][num_envs]  decide (resets of truncated envs), update -> observe
    int32_t *reset_list;     // [num_envs]         screen -> reset (envs truncated by a step the screen kernel settled)
    int *counters;           // [CNT_N]
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
    const DomainProto *t_proto;
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
