// C ABI of libmeshgen_b200.so (include/meshgen_b200.h): handle management, device memory,
// kernel launches.  No torch types, no exceptions across the boundary, no CPU fallback.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "mg_kernels.cu"

using namespace mg;

struct mg_env_s {
    int device = 0;
    int num_envs = 0;
    int max_verts = 0;
    Params P{};
    // owned device buffers
    double2 *t_xy = nullptr;
    double *t_key = nullptr;
    int32_t *t_stamp = nullptr;
    DomainProto *t_proto = nullptr;
    float *t_obs = nullptr;
    mg_episode_stats *d_stats_out = nullptr;
    double2 *sc_tab = nullptr;      // [2][ANGLE_TAB_N] host-libm {sin, cos} of the quantised angles and their halves
    double2 *excl = nullptr;        // [num_envs][cap] not-valid points of mg_move (allocated by its first call)
    int32_t *excl_id = nullptr, *last_excl_id = nullptr;   // their vertex ids; last_not_valid_points (E:570-578)
    bool smooth_pave = true;        // mg_move runs smooth_pave where the reference does (mg_set_option "smooth_pave")
    unsigned char *smooth_scratch = nullptr;
    size_t smooth_slots = 0;
    int32_t *smooth_list = nullptr;   // [num_envs] device list of the envs to smooth
    uint8_t *h_flags = nullptr;       // pinned [num_envs]
    // staging buffers for mg_step_host (used for every caller buffer that is not pinned)
    float *d_act = nullptr, *d_obs = nullptr, *d_term_obs = nullptr;
    double *d_rew = nullptr;
    uint8_t *d_term = nullptr, *d_trunc = nullptr;
    int32_t *d_nel = nullptr;
    cudaStream_t host_stream = nullptr;
    cudaStream_t side_stream = nullptr;           // mg_step_reset_kernel runs here, next to the update kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int32_t *h_cnt = nullptr;       // pinned copy of the step counters (byte accounting of mg_step_host)
    // observation delta (mg_set_obs_delta / mg_set_host_delta): the buffer that is known to hold every env's current
    // observation (written in full by mg_reset or by the previous mg_step with the same pointer)
    bool obs_delta = false;
    const float *obs_bound = nullptr;
    const void *res_bound[4] = {nullptr, nullptr, nullptr, nullptr};   // result delta: the caller arrays the shadows mirror
    bool acct_res_delta = false;
    int64_t last_h2d = 0, last_d2h = 0;
    bool acct_valid = false, acct_obs_rows_delta = false, acct_n_elem = false;
    int acct_term_obs = 0;             // 0 none, 1 pinned (finished rows only), 2 staged (all rows)
    // ordering between the caller's stream (mg_reset / mg_step / mg_snapshot_*) and the private stream of mg_step_host
    cudaStream_t last_user_stream = nullptr;
    bool user_work_pending = false;
    bool ready = false;       // domains or generator configured
    bool was_reset = false;
    int64_t launches = 0;
    // per-kernel timing (mg_set_kernel_timing): CUDA events around the two step kernels
    bool timing = false;
    static constexpr int TIMING_SLOTS = 256;
    cudaEvent_t ev[TIMING_SLOTS][5] = {};
    int64_t timing_steps = 0;
    int sm_count = 148;
    size_t smem = 0, smem_nq = 0;      // one-warp block with / without the integer scratch queue
    int blocks_decide = 16, blocks_update = 16, blocks_observe = 16, blocks_reset = 16;
    bool fuse_decide = true;           // one launch for the decide and update work (mg_set_option "fuse_decide")
    bool reset_side = true;            // resets on the side stream, next to the update kernel (mg_set_option "reset_side")
    bool pdl = true;                   // update / observe kernels as programmatic dependents of their predecessor (mg_set_option "pdl")
    // mg_step_host replays the launches of a step from a CUDA graph (mg_set_option "host_graph"): one graph per distinct
    // (buffers, parameters, options) combination, instantiated the second time the combination is seen
    bool host_graph = true;
    bool host_step_pending = false;    // between mg_step_host_begin and mg_step_host_end
    struct StepGraph {
        StepIO io;
        Params P;
        int opts[8];
        cudaGraphExec_t exec;          // nullptr: seen once, launched kernel by kernel
        int launches;
        uint64_t last_used;
    };
    std::vector<StepGraph> step_graphs;
    uint64_t step_graph_clock = 0;
    std::string err;
};

namespace {

thread_local std::string g_err;

int fail(mg_handle h, int code, const std::string &msg) {
    if (h) h->err = msg;
    g_err = msg;
    return code;
}

#define MG_CUDA(h, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(h, MG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

// Every entry point works on the handle's device and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) {
            status = cudaSetDevice(device);
            switched = status == cudaSuccess;
        }
    }
    ~DeviceGuard() {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
};
#define MG_DEVICE(h)                \
    DeviceGuard guard_((h)->device); \
    MG_CUDA(h, guard_.status)

template <class T>
cudaError_t dalloc(T **p, size_t count) {
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, count * sizeof(T));
    return e;
}

int configure_kernels(mg_handle h) {
    h->smem = smem_bytes(h->P.cap, true);
    h->smem_nq = smem_bytes(h->P.cap, false);
    if (h->smem > 48 * 1024) {
        MG_CUDA(h, cudaFuncSetAttribute(mg_step_observe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        MG_CUDA(h, cudaFuncSetAttribute(mg_step_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        MG_CUDA(h, cudaFuncSetAttribute(mg_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        MG_CUDA(h, cudaFuncSetAttribute(mg_template_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        MG_CUDA(h, cudaFuncSetAttribute(mg_regen_polygon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    }
    if (h->smem_nq > 48 * 1024) {
        MG_CUDA(h, cudaFuncSetAttribute(mg_step_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_nq));
        MG_CUDA(h, cudaFuncSetAttribute(mg_step_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_nq));
    }
    // the vertex rings want shared memory, not L1: ask for the largest carve-out so that the number of
    // resident warps is set by registers, not by the driver's default split
    MG_CUDA(h, cudaFuncSetAttribute(mg_step_decide_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    MG_CUDA(h, cudaFuncSetAttribute(mg_step_update_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    MG_CUDA(h, cudaFuncSetAttribute(mg_step_observe_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    MG_CUDA(h, cudaFuncSetAttribute(mg_step_reset_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // the screen kernel stages its block's records in 16.5 KB of shared memory, four blocks per SM (registers)
    MG_CUDA(h, cudaFuncSetAttribute(mg_step_screen_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 40));
    // resident one-warp blocks per SM of each item kernel (grid = that many blocks: items beyond the first wave are
    // handed out by ticket)
    int nb = 0;
    MG_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mg_step_decide_kernel, 32, h->smem_nq));
    h->blocks_decide = nb > 0 ? nb : 16;
    MG_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mg_step_update_kernel, 32, h->smem_nq));
    h->blocks_update = nb > 0 ? nb : 16;
    MG_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mg_step_observe_kernel, 32, h->smem));
    h->blocks_observe = nb > 0 ? nb : 16;
    h->blocks_reset = 4;               // ~0.3 % of the envs are reset per step: a few one-warp blocks per SM, tickets beyond
    return MG_OK;
}

// sin / cos of the quantised angles with the host's libm -- the library CPython's math.sin / math.cos call in the
// reference (C:154-168, C:946-947, C:1243).  Called through volatile pointers so that the compiler neither folds
// them nor merges the pair into sincos().
int upload_angle_table(mg_handle h) {
    double (*volatile fsin)(double) = ::sin;
    double (*volatile fcos)(double) = ::cos;
    std::vector<double2> tab((size_t)2 * ANGLE_TAB_N);
    for (int k = 0; k < ANGLE_TAB_N; k++) {
        const double a = (double)k / 1e4;             // == round(theta, 4) for every quantised angle
        tab[k] = make_double2(fsin(a), fcos(a));
        tab[(size_t)ANGLE_TAB_N + k] = make_double2(fsin(a / 2), fcos(a / 2));
    }
    MG_CUDA(h, cudaMemcpy(h->sc_tab, tab.data(), tab.size() * sizeof(double2), cudaMemcpyHostToDevice));
    h->P.sc_full = h->sc_tab;
    h->P.sc_half = h->sc_tab + ANGLE_TAB_N;
    return MG_OK;
}

// Device alias of a caller's host buffer when it is pinned / registered (torch .pin_memory(), cudaHostAlloc,
// cudaHostRegister); nullptr for pageable memory.  Queried on every call (about a microsecond): a cached answer
// would go stale if the caller freed the buffer and another allocation reused the address.
template <class T>
T *pinned_alias(T *host) {
    if (!host) return nullptr;
    cudaPointerAttributes a{};
    void *dev = nullptr;
    if (cudaPointerGetAttributes(&a, host) == cudaSuccess && a.type == cudaMemoryTypeHost) dev = a.devicePointer;
    cudaGetLastError();
    return (T *)dev;
}

void free_templates(mg_handle h) {
    cudaFree(h->t_xy); cudaFree(h->t_key); cudaFree(h->t_stamp); cudaFree(h->t_proto); cudaFree(h->t_obs);
    h->t_xy = nullptr; h->t_key = nullptr; h->t_stamp = nullptr; h->t_proto = nullptr; h->t_obs = nullptr;
    h->P.t_xy = nullptr; h->P.t_key = nullptr; h->P.t_stamp = nullptr; h->P.t_proto = nullptr; h->P.t_obs = nullptr;
    h->P.n_domains = 0;
}

void note_user_stream(mg_handle h, cudaStream_t s) {
    h->last_user_stream = s;
    h->user_work_pending = true;
}

// The launches of one step on stream s: screen, [decide,] update, observe -- and, with auto-reset, the reset kernel on
// the side stream between the screen kernel and the end of the step (event fork / join: capturable in a CUDA graph).
// The side stream has the default priority on purpose: with a higher one the resets take the first slots that free up
// (the failed decisions, ~7 us into the update kernel) and slow the update kernel's busy phase down by 7 us; as it is
// they start in its tail (profiles/README.md, round 2).
int launch_step(mg_handle h, const StepIO &io, cudaStream_t s) {
    const int N = h->num_envs;
    auto grid = [&](int per_sm) { const int r = h->sm_count * per_sm; return N < r ? N : r; };
    cudaEvent_t *ev = h->timing ? h->ev[h->timing_steps % mg_env_s::TIMING_SLOTS] : nullptr;
    const bool side = h->P.auto_reset != 0 && h->reset_side;
    if (ev) cudaEventRecord(ev[0], s);
    mg_step_screen_kernel<<<(N + SCREEN_THREADS - 1) / SCREEN_THREADS, SCREEN_THREADS, 0, s>>>(h->P, io);
    if (ev) cudaEventRecord(ev[1], s);
    if (side) {
        MG_CUDA(h, cudaEventRecord(h->ev_fork, s));
        MG_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        mg_step_reset_kernel<<<grid(h->blocks_reset), 32, h->smem, h->side_stream>>>(h->P, io);
        MG_CUDA(h, cudaEventRecord(h->ev_join, h->side_stream));
    }
    if (!h->fuse_decide) mg_step_decide_kernel<<<grid(h->blocks_decide), 32, h->smem_nq, s>>>(h->P, io);
    if (ev) cudaEventRecord(ev[2], s);
    const bool pdl = h->pdl && !ev;
    auto launch_dependent = [&](auto kernel, int blocks, size_t smem, auto... args) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)blocks); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, args...);
    };
    if (pdl && h->fuse_decide) MG_CUDA(h, launch_dependent(mg_step_update_kernel, grid(h->blocks_update), h->smem_nq, h->P, io, 1));
    else mg_step_update_kernel<<<grid(h->blocks_update), 32, h->smem_nq, s>>>(h->P, io, h->fuse_decide ? 1 : 0);
    if (ev) cudaEventRecord(ev[3], s);
    if (h->P.auto_reset != 0 && !side) mg_step_reset_kernel<<<grid(h->blocks_reset), 32, h->smem, s>>>(h->P, io);
    if (pdl && (side || h->P.auto_reset == 0)) MG_CUDA(h, launch_dependent(mg_step_observe_kernel, grid(h->blocks_observe), h->smem, h->P, io));
    else mg_step_observe_kernel<<<grid(h->blocks_observe), 32, h->smem, s>>>(h->P, io);
    if (side) MG_CUDA(h, cudaStreamWaitEvent(s, h->ev_join, 0));
    if (ev) { cudaEventRecord(ev[4], s); h->timing_steps++; }
    h->launches += (h->fuse_decide ? 3 : 4) + (h->P.auto_reset != 0 ? 1 : 0);
    MG_CUDA(h, cudaGetLastError());
    return MG_OK;
}

void drop_step_graphs(mg_handle h) {
    for (auto &g : h->step_graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->step_graphs.clear();
}

// launch_step for mg_step_host: the same launches replayed from a CUDA graph -- one cudaGraphLaunch instead of four or five
// kernel launches, two event records and two stream waits, and the kernels of a step then follow each other on the device
// without the gaps of separate launches (profiles/README.md: a graph-replayed step is ~10 % shorter than an eagerly
// launched one).  A graph is keyed by everything its kernels were launched with (StepIO, Params, options), built the
// second time the key is seen (a caller that passes fresh buffers on every call never pays for a capture), and the
// cache is bounded.
int launch_step_cached(mg_handle h, const StepIO &io, cudaStream_t s) {
    if (!h->host_graph || h->timing) return launch_step(h, io, s);
    constexpr size_t MAX_GRAPHS = 96;
    const int opts[8] = {h->fuse_decide, h->reset_side, h->pdl, h->blocks_decide, h->blocks_update, h->blocks_observe, h->blocks_reset, 0};
    mg_env_s::StepGraph *hit = nullptr;
    for (auto &g : h->step_graphs)
        if (std::memcmp(&g.io, &io, sizeof(StepIO)) == 0 && std::memcmp(&g.P, &h->P, sizeof(Params)) == 0 &&
            std::memcmp(g.opts, opts, sizeof(opts)) == 0) { hit = &g; break; }
    if (!hit) {
        if (h->step_graphs.size() >= MAX_GRAPHS) {              // evict the entry that was used longest ago
            size_t victim = 0;
            for (size_t k = 1; k < h->step_graphs.size(); k++)
                if (h->step_graphs[k].last_used < h->step_graphs[victim].last_used) victim = k;
            if (h->step_graphs[victim].exec) cudaGraphExecDestroy(h->step_graphs[victim].exec);
            h->step_graphs.erase(h->step_graphs.begin() + victim);
        }
        mg_env_s::StepGraph g;
        std::memcpy(&g.io, &io, sizeof(StepIO)); std::memcpy(&g.P, &h->P, sizeof(Params)); std::memcpy(g.opts, opts, sizeof(opts));
        g.exec = nullptr; g.launches = 0; g.last_used = ++h->step_graph_clock;
        h->step_graphs.push_back(g);
        return launch_step(h, io, s);
    }
    hit->last_used = ++h->step_graph_clock;
    if (!hit->exec) {
        MG_CUDA(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const int64_t l0 = h->launches;
        const int rc = launch_step(h, io, s);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(s, &graph);
        hit->launches = (int)(h->launches - l0);
        h->launches = l0;
        if (rc != MG_OK || e != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            h->host_graph = false;                              // (capture not possible here: stay on plain launches)
            drop_step_graphs(h);
            return launch_step(h, io, s);
        }
        const cudaError_t ei = cudaGraphInstantiate(&hit->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            cudaGetLastError();
            h->host_graph = false;
            drop_step_graphs(h);
            return launch_step(h, io, s);
        }
    }
    MG_CUDA(h, cudaGraphLaunch(hit->exec, s));
    h->launches += hit->launches;
    return MG_OK;
}

// Random-polygon mode: regenerate the polygon of (env, episode) -- episode < 0 = the env's current one -- and copy up
// to max_vertices of it to the host (see mg_regen_polygon_kernel).  Synchronises.
int regen_polygon(mg_handle h, int env, int episode, double *xy_host, int max_vertices, int32_t *n_out, double *area_out,
                  int32_t *coarse_px_host, int32_t *k_out, double *spacing_out) {
    const int cap = h->P.cap;
    char *buf = nullptr;
    const size_t off_n = sizeof(double2) * cap, off_dbg = off_n + 16, off_area = off_dbg + sizeof(double) * 66;
    MG_CUDA(h, cudaMalloc((void **)&buf, off_area + 16));
    mg_regen_polygon_kernel<<<1, 32, h->smem>>>(h->P, env, episode, (double2 *)buf, (int32_t *)(buf + off_n), (double *)(buf + off_dbg),
                                                (double *)(buf + off_area));
    h->launches++;
    int32_t n = 0;
    double dbg[66], area = 0;
    cudaError_t e = cudaMemcpy(&n, buf + off_n, sizeof(n), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dbg, buf + off_dbg, sizeof(dbg), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&area, buf + off_area, sizeof(area), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && xy_host && n > 0 && n <= cap) {
        const int c = n < max_vertices ? n : max_vertices;
        if (c > 0) e = cudaMemcpy(xy_host, buf, sizeof(double2) * c, cudaMemcpyDeviceToHost);
    }
    cudaFree(buf);
    if (e != cudaSuccess) return fail(h, MG_ERR_CUDA, std::string("mg_regen_polygon_kernel: ") + cudaGetErrorString(e));
    if (n_out) *n_out = n;
    if (area_out) *area_out = area;
    const int K = (int)dbg[0];
    if (k_out) *k_out = K;
    if (spacing_out) *spacing_out = dbg[1];
    if (coarse_px_host)
        for (int i = 0; i < 2 * K && i < 64; i++) coarse_px_host[i] = (int32_t)dbg[2 + i];
    return MG_OK;
}

}  // namespace

extern "C" {

const char *mg_version(void) { return "meshgen_b200 0.2 (sm_100a)"; }

const char *mg_last_error(mg_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int mg_create(mg_handle *out, int device, int num_envs, int max_verts) {
    if (!out || num_envs <= 0 || max_verts < 4 || max_verts > 8192) return fail(nullptr, MG_ERR_ARG, "mg_create: bad argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MG_ERR_CUDA, std::string("mg_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, MG_ERR_ARG, "mg_create: bad device index");
    mg_handle h = new (std::nothrow) mg_env_s();
    if (!h) return fail(nullptr, MG_ERR_ARG, "mg_create: out of host memory");
    h->device = device; h->num_envs = num_envs; h->max_verts = max_verts;
    DeviceGuard guard(device);
    if (guard.status != cudaSuccess) {
        fail(nullptr, MG_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(guard.status));
        delete h;
        return MG_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    Params &P = h->P;
    P.num_envs = num_envs;
    P.auto_reset = 1;
    P.cap = (max_verts + 1) & ~1;
    const size_t NC = (size_t)num_envs * P.cap;
    // per-env element / inserted-vertex logs: 2 x max_verts entries each by default (what a random or half-trained
    // policy produces before it is truncated); evaluation of a trained policy on a large domain needs more
    // (element counts scale with the domain's area: mg_set_log_capacity).  The element COUNT is always exact.
    P.elem_cap = 2 * P.cap < 64 ? 64 : 2 * P.cap;
    P.ins_cap = P.elem_cap;
    int rc = MG_OK;
    auto A = [&](cudaError_t er, const char *what) {
        if (er != cudaSuccess && rc == MG_OK) rc = fail(h, MG_ERR_CUDA, std::string("cudaMalloc ") + what + ": " + cudaGetErrorString(er));
    };
    A(dalloc(&P.xy, NC), "xy"); A(dalloc(&P.key, NC), "key"); A(dalloc(&P.stamp, NC), "stamp"); A(dalloc(&P.vid, NC), "vid");
    A(dalloc(&P.hot, (size_t)num_envs), "hot"); A(dalloc(&P.cold, (size_t)num_envs), "cold");
    A(dalloc(&P.stats, (size_t)STAT_SLOTS), "stats");
    A(dalloc(&P.obs_cache, (size_t)num_envs * MG_OBS_DIM), "obs");
    A(dalloc(&P.decide_list, (size_t)DECIDE_SEGS * num_envs), "decide_list"); A(dalloc(&P.accept_list, (size_t)NBINS * num_envs), "accept_list");
    A(dalloc(&P.observe_list, (size_t)NBINS * num_envs), "observe_list"); A(dalloc(&P.counters, (size_t)CNT_N), "counters");
    A(dalloc(&P.reset_list, (size_t)num_envs), "reset_list");
    A(dalloc(&P.elem, (size_t)num_envs * P.elem_cap * 4), "elem"); A(dalloc(&P.ins_xy, (size_t)num_envs * P.ins_cap), "ins_xy");
    A(dalloc(&h->d_stats_out, 1), "stats_out");
    A(dalloc(&h->sc_tab, (size_t)2 * ANGLE_TAB_N), "angle table");
    A(dalloc(&h->d_act, (size_t)num_envs * 3), "act"); A(dalloc(&h->d_obs, (size_t)num_envs * MG_OBS_DIM), "obs_out");
    A(dalloc(&h->d_term_obs, (size_t)num_envs * MG_OBS_DIM), "term_obs"); A(dalloc(&h->d_rew, (size_t)num_envs), "rew");
    A(dalloc(&h->d_term, (size_t)num_envs), "term"); A(dalloc(&h->d_trunc, (size_t)num_envs), "trunc");
    A(dalloc(&h->d_nel, (size_t)num_envs), "nel");
    A(cudaMallocHost((void **)&h->h_cnt, CNT_N * sizeof(int32_t)), "h_cnt");
    if (rc == MG_OK && (cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking) != cudaSuccess ||
                        cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
                        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess))
        rc = fail(h, MG_ERR_CUDA, "cudaStreamCreate / cudaEventCreate");
    if (rc == MG_OK) rc = configure_kernels(h);
    if (rc == MG_OK) rc = upload_angle_table(h);
    if (rc != MG_OK) { g_err = h->err; mg_destroy(h); return rc; }
    *out = h;
    return MG_OK;
}

int mg_destroy(mg_handle h) {
    if (!h) return MG_OK;
    DeviceGuard guard(h->device);
    Params &P = h->P;
    cudaFree(P.xy); cudaFree(P.key); cudaFree(P.stamp); cudaFree(P.vid); cudaFree(P.hot); cudaFree(P.cold); cudaFree(P.stats);
    cudaFree(P.obs_cache); cudaFree(P.elem); cudaFree(P.ins_xy);
    cudaFree(P.decide_list); cudaFree(P.accept_list); cudaFree(P.observe_list); cudaFree(P.counters); cudaFree(P.reset_list);
    free_templates(h);
    cudaFree(h->sc_tab); cudaFree(h->excl); cudaFree(h->excl_id); cudaFree(h->last_excl_id); cudaFree(h->smooth_scratch);
    cudaFree(h->smooth_list); cudaFreeHost(h->h_flags);
    cudaFree(h->d_stats_out); cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_term_obs); cudaFree(h->d_rew);
    cudaFree(h->d_term); cudaFree(h->d_trunc); cudaFree(h->d_nel);
    cudaFreeHost(h->h_cnt);
    for (auto &tr : h->ev)
        for (cudaEvent_t e : tr)
            if (e) cudaEventDestroy(e);
    drop_step_graphs(h);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return MG_OK;
}

int mg_set_auto_reset(mg_handle h, int enabled) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_set_auto_reset: null handle");
    h->P.auto_reset = enabled ? 1 : 0;
    return MG_OK;
}

int mg_replay_add(mg_handle h, int64_t capacity_steps, int64_t slot, float *buf_obs, float *buf_next_obs, float *buf_act,
                  float *buf_rew, uint8_t *buf_done, uint8_t *buf_timeout, const float *prev_obs, const float *act,
                  const float *new_obs, const double *rew, const uint8_t *term, const uint8_t *trunc,
                  const float *term_obs, void *stream) {
    if (!h || !buf_obs || !buf_next_obs || !buf_act || !buf_rew || !buf_done || !buf_timeout || !prev_obs || !act || !new_obs ||
        !rew || !term || !trunc || !term_obs)
        return fail(h, MG_ERR_ARG, "mg_replay_add: null pointer");
    if (capacity_steps <= 0 || slot < 0 || slot >= capacity_steps) return fail(h, MG_ERR_ARG, "mg_replay_add: slot out of range");
    MG_DEVICE(h);
    const size_t N = (size_t)h->num_envs, s = (size_t)slot;
    const int total = h->num_envs * MG_OBS_DIM;
    mg_replay_add_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        h->num_envs, buf_obs + s * N * MG_OBS_DIM, buf_next_obs + s * N * MG_OBS_DIM, buf_act + s * N * MG_ACT_DIM, buf_rew + s * N,
        buf_done + s * N, buf_timeout + s * N, prev_obs, act, new_obs, rew, term, trunc, term_obs);
    h->launches += 1;
    MG_CUDA(h, cudaGetLastError());
    return MG_OK;
}

// ---- snapshot / restore --------------------------------------------------------------------------
namespace {
struct Piece { void *ptr; size_t bytes; };
std::vector<Piece> snapshot_pieces(mg_handle h) {
    Params &P = h->P;
    const size_t N = (size_t)h->num_envs, NC = N * P.cap;
    return {
        {P.xy, NC * sizeof(double2)}, {P.key, NC * sizeof(double)}, {P.stamp, NC * sizeof(int32_t)}, {P.vid, NC * sizeof(int32_t)},
        {P.hot, N * sizeof(EnvHot)}, {P.cold, N * sizeof(EnvCold)}, {P.stats, STAT_SLOTS * sizeof(StatsAcc)},
        {P.obs_cache, N * MG_OBS_DIM * sizeof(float)},
        {P.elem, N * P.elem_cap * 4 * sizeof(int32_t)}, {P.ins_xy, N * P.ins_cap * sizeof(double2)},
        {P.counters, CNT_N * sizeof(int)},
    };
}
constexpr size_t SNAP_ALIGN = 256;
size_t snap_round(size_t b) { return (b + SNAP_ALIGN - 1) / SNAP_ALIGN * SNAP_ALIGN; }
// everything a blob must agree on with the handle it is loaded into
struct SnapHeader {
    int64_t magic, version, total_bytes;
    int64_t num_envs, cap, random_mode, elem_cap, ins_cap, n_domains;
    uint64_t seed;
    int64_t env_id_offset;
};
static_assert(sizeof(SnapHeader) <= SNAP_ALIGN, "snapshot header");
constexpr int64_t SNAP_MAGIC = 0x4d4753324e415053ll, SNAP_VERSION = 4;
SnapHeader snap_header(mg_handle h, int64_t total) {
    SnapHeader H{};
    H.magic = SNAP_MAGIC; H.version = SNAP_VERSION; H.total_bytes = total;
    H.num_envs = h->num_envs; H.cap = h->P.cap; H.random_mode = h->P.random_mode; H.elem_cap = h->P.elem_cap;
    H.ins_cap = h->P.ins_cap; H.n_domains = h->P.random_mode ? 0 : h->P.n_domains;
    H.seed = h->P.random_mode ? h->P.seed : 0; H.env_id_offset = h->P.random_mode ? h->P.env_id_offset : 0;
    return H;
}
}  // namespace

int64_t mg_snapshot_bytes(mg_handle h) {
    if (!h) return 0;
    size_t total = SNAP_ALIGN;                       // header
    for (const Piece &p : snapshot_pieces(h)) total += snap_round(p.bytes);
    return (int64_t)total;
}

int mg_snapshot_save(mg_handle h, void *blob_dev, void *stream) {
    if (!h || !blob_dev) return fail(h, MG_ERR_ARG, "mg_snapshot_save: null pointer");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_snapshot_save: call mg_reset first");
    if (h->host_step_pending) return fail(h, MG_ERR_STATE, "mg_snapshot_save: a host step is in flight (mg_step_host_end first)");
    MG_DEVICE(h);
    cudaStream_t s = (cudaStream_t)stream;
    const SnapHeader hdr = snap_header(h, mg_snapshot_bytes(h));
    MG_CUDA(h, cudaMemcpyAsync(blob_dev, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, s));
    MG_CUDA(h, cudaStreamSynchronize(s));            // hdr is a stack object
    char *dst = (char *)blob_dev + SNAP_ALIGN;
    for (const Piece &p : snapshot_pieces(h)) {
        MG_CUDA(h, cudaMemcpyAsync(dst, p.ptr, p.bytes, cudaMemcpyDeviceToDevice, s));
        dst += snap_round(p.bytes);
    }
    note_user_stream(h, s);
    return MG_OK;
}

int mg_snapshot_load(mg_handle h, const void *blob_dev, int64_t blob_bytes, void *stream) {
    if (!h || !blob_dev) return fail(h, MG_ERR_ARG, "mg_snapshot_load: null pointer");
    if (!h->ready) return fail(h, MG_ERR_STATE, "mg_snapshot_load: configure domains or the generator first");
    if (blob_bytes < (int64_t)SNAP_ALIGN) return fail(h, MG_ERR_ARG, "mg_snapshot_load: blob shorter than its header");
    if (h->host_step_pending) return fail(h, MG_ERR_STATE, "mg_snapshot_load: a host step is in flight (mg_step_host_end first)");
    MG_DEVICE(h);
    cudaStream_t s = (cudaStream_t)stream;
    SnapHeader got{};
    MG_CUDA(h, cudaMemcpyAsync(&got, blob_dev, sizeof(got), cudaMemcpyDeviceToHost, s));
    MG_CUDA(h, cudaStreamSynchronize(s));
    const SnapHeader want = snap_header(h, mg_snapshot_bytes(h));
    if (got.magic != want.magic || got.version != want.version)
        return fail(h, MG_ERR_ARG, "mg_snapshot_load: not a snapshot of this library version");
    if (got.total_bytes != want.total_bytes || blob_bytes < want.total_bytes)
        return fail(h, MG_ERR_ARG, "mg_snapshot_load: blob size does not match this handle (truncated blob or different capacities)");
    if (got.num_envs != want.num_envs || got.cap != want.cap || got.random_mode != want.random_mode || got.elem_cap != want.elem_cap ||
        got.ins_cap != want.ins_cap || got.n_domains != want.n_domains || got.seed != want.seed || got.env_id_offset != want.env_id_offset)
        return fail(h, MG_ERR_ARG, "mg_snapshot_load: blob does not match this handle (num_envs / max_verts / mode / log capacity / "
                                   "domains / generator seed / env id offset)");
    const char *src = (const char *)blob_dev + SNAP_ALIGN;
    for (const Piece &p : snapshot_pieces(h)) {
        MG_CUDA(h, cudaMemcpyAsync(p.ptr, src, p.bytes, cudaMemcpyDeviceToDevice, s));
        src += snap_round(p.bytes);
    }
    h->was_reset = true;
    h->obs_bound = nullptr;              // no caller buffer holds the restored observations yet
    note_user_stream(h, s);
    return MG_OK;
}

int mg_num_envs(mg_handle h) { return h ? h->num_envs : 0; }
int mg_max_verts(mg_handle h) { return h ? h->max_verts : 0; }
int64_t mg_launch_count(mg_handle h) { return h ? h->launches : 0; }

int mg_set_domains(mg_handle h, const double *xy_host, const int32_t *offsets_host, int n_domains,
                   const int32_t *env_domain_host, const double *areas_host) {
    if (!h || !xy_host || !offsets_host || !env_domain_host || n_domains <= 0) return fail(h, MG_ERR_ARG, "mg_set_domains: bad argument");
    MG_DEVICE(h);
    Params &P = h->P;
    const int cap = P.cap;
    std::vector<double2> txy((size_t)n_domains * cap, make_double2(0, 0));
    std::vector<int32_t> n0s(n_domains);
    for (int d = 0; d < n_domains; d++) {
        int n = offsets_host[d + 1] - offsets_host[d];
        if (n < 3) return fail(h, MG_ERR_ARG, "mg_set_domains: polygon with fewer than 3 vertices");
        if (n > h->max_verts) return fail(h, MG_ERR_CAPACITY, "mg_set_domains: polygon larger than max_verts");
        for (int j = 0; j < n; j++) {
            const double *p = xy_host + 2 * ((size_t)offsets_host[d] + j);
            txy[(size_t)d * cap + j] = make_double2(p[0], p[1]);
        }
        n0s[d] = n;
    }
    std::vector<EnvCold> cold(h->num_envs);
    std::memset(cold.data(), 0, sizeof(EnvCold) * cold.size());
    for (int e = 0; e < h->num_envs; e++) {
        int d = env_domain_host[e];
        if (d < 0 || d >= n_domains) return fail(h, MG_ERR_ARG, "mg_set_domains: env_domain out of range");
        cold[e].domain = d;
    }
    MG_CUDA(h, cudaDeviceSynchronize());
    h->ready = false; h->was_reset = false;          // a failure below leaves the handle unconfigured, not half configured
    drop_step_graphs(h);                             // (they were captured with the old templates)
    free_templates(h);
    MG_CUDA(h, dalloc(&h->t_xy, (size_t)n_domains * cap));
    MG_CUDA(h, dalloc(&h->t_key, (size_t)n_domains * cap));
    MG_CUDA(h, dalloc(&h->t_stamp, (size_t)n_domains * cap));
    MG_CUDA(h, dalloc(&h->t_proto, (size_t)n_domains));
    MG_CUDA(h, dalloc(&h->t_obs, (size_t)n_domains * MG_OBS_DIM));
    MG_CUDA(h, cudaMemcpy(h->t_xy, txy.data(), sizeof(double2) * txy.size(), cudaMemcpyHostToDevice));
    MG_CUDA(h, cudaMemcpy(P.cold, cold.data(), sizeof(EnvCold) * cold.size(), cudaMemcpyHostToDevice));
    MG_CUDA(h, cudaMemset(P.hot, 0, sizeof(EnvHot) * h->num_envs));
    double *d_areas = nullptr;
    int32_t *d_n0 = nullptr;
    MG_CUDA(h, cudaMalloc((void **)&d_n0, sizeof(int32_t) * n_domains));
    MG_CUDA(h, cudaMemcpy(d_n0, n0s.data(), sizeof(int32_t) * n_domains, cudaMemcpyHostToDevice));
    if (areas_host) {
        MG_CUDA(h, cudaMalloc((void **)&d_areas, sizeof(double) * n_domains));
        MG_CUDA(h, cudaMemcpy(d_areas, areas_host, sizeof(double) * n_domains, cudaMemcpyHostToDevice));
    }
    P.n_domains = n_domains; P.random_mode = 0;
    P.t_xy = h->t_xy; P.t_key = h->t_key; P.t_stamp = h->t_stamp; P.t_proto = h->t_proto; P.t_obs = h->t_obs;
    mg_template_kernel<<<n_domains, 32, h->smem>>>(P, h->t_xy, h->t_key, h->t_stamp, h->t_proto, h->t_obs, d_n0, d_areas);
    h->launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (d_areas) cudaFree(d_areas);
    cudaFree(d_n0);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, MG_ERR_CUDA, std::string("mg_template_kernel: ") + cudaGetErrorString(e));
    h->ready = true;
    return MG_OK;
}

int mg_set_random(mg_handle h, uint64_t seed, const mg_polygen_cfg *cfg, int64_t env_id_offset) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_set_random: null handle");
    mg_polygen_cfg c;
    if (cfg) c = *cfg;
    else {
        c.ctr_x = 250; c.ctr_y = 250; c.ave_radius = 100; c.irregularity = 0.55; c.spikeyness = 0.7;
        c.min_coarse = 8; c.max_coarse = 24; c.min_verts = 64; c.max_verts = 512;
    }
    if (c.min_coarse < 3 || c.max_coarse < c.min_coarse || c.max_coarse > 32 || c.min_verts < 8 || c.max_verts < c.min_verts ||
        c.max_verts > h->max_verts || c.min_verts < c.max_coarse)
        return fail(h, MG_ERR_ARG, "mg_set_random: bad generator configuration (need 3<=min_coarse<=max_coarse<=32, "
                                   "max_coarse<=min_verts<=max_verts<=handle max_verts)");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    MG_CUDA(h, cudaMemset(h->P.hot, 0, sizeof(EnvHot) * h->num_envs));
    MG_CUDA(h, cudaMemset(h->P.cold, 0, sizeof(EnvCold) * h->num_envs));
    h->P.random_mode = 1; h->P.seed = seed; h->P.gen = c; h->P.env_id_offset = env_id_offset;
    drop_step_graphs(h);
    h->ready = true; h->was_reset = false;
    return MG_OK;
}

int mg_reset(mg_handle h, const uint8_t *mask_dev, float *obs_dev, void *stream) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_reset: null handle");
    if (!h->ready) return fail(h, MG_ERR_STATE, "mg_reset: call mg_set_domains or mg_set_random first");
    if (mask_dev && !h->was_reset) return fail(h, MG_ERR_STATE, "mg_reset: the first reset must cover all envs (mask = NULL)");
    if (h->host_step_pending) return fail(h, MG_ERR_STATE, "mg_reset: a host step is in flight (mg_step_host_end first)");
    MG_DEVICE(h);
    mg_reset_kernel<<<h->num_envs, 32, h->smem, (cudaStream_t)stream>>>(h->P, mask_dev, obs_dev);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    h->was_reset = true;
    h->obs_bound = obs_dev;              // every row was written (NULL: no caller buffer holds the observations)
    note_user_stream(h, (cudaStream_t)stream);
    return MG_OK;
}

int mg_step(mg_handle h, const float *act_dev, float *obs_dev, double *rew_dev, uint8_t *term_dev, uint8_t *trunc_dev,
            float *term_obs_dev, int32_t *n_elem_dev, void *stream) {
    if (!h || !act_dev || !obs_dev || !rew_dev || !term_dev || !trunc_dev) return fail(h, MG_ERR_ARG, "mg_step: null pointer");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_step: call mg_reset first");
    if (h->host_step_pending) return fail(h, MG_ERR_STATE, "mg_step: a host step is in flight (mg_step_host_end first)");
    MG_DEVICE(h);
    StepIO io;
    io.act = act_dev; io.obs_out = obs_dev; io.rew_out = rew_dev; io.term_out = term_dev; io.trunc_out = trunc_dev;
    io.term_obs_out = term_obs_dev; io.n_elem_out = n_elem_dev;
    io.rew_sh = nullptr; io.term_sh = nullptr; io.trunc_sh = nullptr; io.nel_sh = nullptr; io.res_full = 1;
    io.obs_full = (h->obs_delta && h->obs_bound == obs_dev) ? 0 : 1;
    const int rc = launch_step(h, io, (cudaStream_t)stream);
    if (rc != MG_OK) return rc;
    h->obs_bound = obs_dev;
    note_user_stream(h, (cudaStream_t)stream);
    return MG_OK;
}

int mg_step_host(mg_handle h, const float *act_host, float *obs_host, double *rew_host, uint8_t *term_host,
                 uint8_t *trunc_host, float *term_obs_host, int32_t *n_elem_host) {
    const int rc = mg_step_host_begin(h, act_host, obs_host, rew_host, term_host, trunc_host, term_obs_host, n_elem_host);
    return rc != MG_OK ? rc : mg_step_host_end(h);
}

int mg_step_host_end(mg_handle h) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_step_host_end: null handle");
    if (!h->host_step_pending) return fail(h, MG_ERR_STATE, "mg_step_host_end: no mg_step_host_begin in flight");
    MG_DEVICE(h);
    h->host_step_pending = false;
    MG_CUDA(h, cudaStreamSynchronize(h->host_stream));
    h->acct_valid = true;      // (mg_last_host_bytes reads the step counters only when asked)
    return MG_OK;
}

int mg_step_host_begin(mg_handle h, const float *act_host, float *obs_host, double *rew_host, uint8_t *term_host,
                       uint8_t *trunc_host, float *term_obs_host, int32_t *n_elem_host) {
    if (!h || !act_host || !obs_host || !rew_host || !term_host || !trunc_host) return fail(h, MG_ERR_ARG, "mg_step_host: null pointer");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_step_host: call mg_reset first");
    if (h->host_step_pending) return fail(h, MG_ERR_STATE, "mg_step_host_begin: the previous host step has not been ended");
    MG_DEVICE(h);
    const size_t N = h->num_envs;
    cudaStream_t s = h->host_stream;
    // work the caller enqueued on its own stream (mg_reset, mg_step, mg_snapshot_*) must finish before this step
    // touches the env state on the library's private stream
    if (h->user_work_pending) {
        MG_CUDA(h, cudaStreamSynchronize(h->last_user_stream));
        h->user_work_pending = false;
    }
    // Pinned caller buffers are used by the step kernels themselves through their device aliases: the screen kernel
    // reads the actions over PCIe while it loads the env records (no copy in front of the step), results are posted
    // PCIe writes that overlap the later kernels (no staging copy, no copy-engine round after the step).  Pageable
    // buffers go through the handle's staging buffers and one cudaMemcpyAsync each.
    const float *act_a = pinned_alias(act_host);
    if (!act_a) MG_CUDA(h, cudaMemcpyAsync(h->d_act, act_host, N * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    float *obs_a = pinned_alias(obs_host), *tobs_a = pinned_alias(term_obs_host);
    double *rew_a = pinned_alias(rew_host);
    uint8_t *term_a = pinned_alias(term_host), *trunc_a = pinned_alias(trunc_host);
    int32_t *nel_a = pinned_alias(n_elem_host);
    StepIO io;
    std::memset(&io, 0, sizeof(io));                            // (compared bytewise by launch_step_cached)
    io.act = act_a ? act_a : h->d_act;
    io.obs_out = obs_a ? obs_a : h->d_obs;
    io.rew_out = rew_a ? rew_a : h->d_rew;
    io.term_out = term_a ? term_a : h->d_term;
    io.trunc_out = trunc_a ? trunc_a : h->d_trunc;
    io.term_obs_out = term_obs_host ? (tobs_a ? tobs_a : h->d_term_obs) : nullptr;
    io.n_elem_out = n_elem_host ? (nel_a ? nel_a : h->d_nel) : nullptr;
    io.obs_full = (h->obs_delta && h->obs_bound == io.obs_out) ? 0 : 1;
    // Result delta: with pinned reward / flag / count arrays in delta mode the kernels write a value over PCIe only where it
    // differs from what the caller's array already holds (the handle's staging buffers keep a device copy of that).
    io.rew_sh = nullptr; io.term_sh = nullptr; io.trunc_sh = nullptr; io.nel_sh = nullptr; io.res_full = 1;
    const bool res_delta = h->obs_delta && rew_a && term_a && trunc_a && (!n_elem_host || nel_a);
    if (res_delta) {
        io.rew_sh = h->d_rew; io.term_sh = h->d_term; io.trunc_sh = h->d_trunc; io.nel_sh = n_elem_host ? h->d_nel : nullptr;
        io.res_full = (h->res_bound[0] == rew_a && h->res_bound[1] == term_a && h->res_bound[2] == trunc_a && h->res_bound[3] == nel_a) ? 0 : 1;
    }
    const int rc = launch_step_cached(h, io, s);
    if (rc != MG_OK) return rc;
    h->obs_bound = io.obs_out;
    h->res_bound[0] = res_delta ? rew_a : nullptr; h->res_bound[1] = res_delta ? term_a : nullptr;
    h->res_bound[2] = res_delta ? trunc_a : nullptr; h->res_bound[3] = res_delta ? nel_a : nullptr;
    h->acct_res_delta = res_delta && !io.res_full;
    if (!obs_a) MG_CUDA(h, cudaMemcpyAsync(obs_host, h->d_obs, N * MG_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (!rew_a) MG_CUDA(h, cudaMemcpyAsync(rew_host, h->d_rew, N * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (!term_a) MG_CUDA(h, cudaMemcpyAsync(term_host, h->d_term, N, cudaMemcpyDeviceToHost, s));
    if (!trunc_a) MG_CUDA(h, cudaMemcpyAsync(trunc_host, h->d_trunc, N, cudaMemcpyDeviceToHost, s));
    if (term_obs_host && !tobs_a)
        MG_CUDA(h, cudaMemcpyAsync(term_obs_host, h->d_term_obs, N * MG_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (n_elem_host && !nel_a) MG_CUDA(h, cudaMemcpyAsync(n_elem_host, h->d_nel, N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    // what the byte accounting of this step needs
    h->host_step_pending = true;
    h->acct_valid = false;
    h->acct_obs_rows_delta = obs_a && !io.obs_full;
    h->acct_term_obs = term_obs_host ? (tobs_a ? 1 : 2) : 0;
    h->acct_n_elem = n_elem_host != nullptr;
    return MG_OK;
}

int mg_move(mg_handle h, const double *polar_dev, const double *type_dev, float *obs_dev, uint8_t *done_dev, uint8_t *complete_dev,
            uint8_t *exhausted_dev, int32_t *n_elem_dev, void *stream) {
    if (!h || !polar_dev || !type_dev || !obs_dev || !done_dev || !complete_dev || !exhausted_dev)
        return fail(h, MG_ERR_ARG, "mg_move: null pointer");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_move: call mg_reset first");
    if (h->host_step_pending) return fail(h, MG_ERR_STATE, "mg_move: a host step is in flight (mg_step_host_end first)");
    MG_DEVICE(h);
    if (!h->excl) {
        MG_CUDA(h, dalloc(&h->excl, (size_t)h->num_envs * h->P.cap));
        MG_CUDA(h, dalloc(&h->excl_id, (size_t)h->num_envs * h->P.cap));
        MG_CUDA(h, dalloc(&h->last_excl_id, (size_t)h->num_envs * h->P.cap));
        MG_CUDA(h, dalloc(&h->smooth_list, (size_t)h->num_envs));
        MG_CUDA(h, cudaMallocHost((void **)&h->h_flags, (size_t)h->num_envs));
        MG_CUDA(h, cudaFuncSetAttribute(mg_move_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        MG_CUDA(h, cudaFuncSetAttribute(mg_smooth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    }
    cudaStream_t s = (cudaStream_t)stream;
    MoveIO io;
    io.polar = polar_dev; io.type = type_dev; io.obs_out = obs_dev; io.done_out = done_dev; io.complete_out = complete_dev;
    io.exhausted_out = exhausted_dev; io.n_elem_out = n_elem_dev;
    mg_move_kernel<<<h->num_envs, 32, h->smem, s>>>(h->P, io, h->excl, h->excl_id);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    h->obs_bound = obs_dev;              // every row was written
    note_user_stream(h, s);
    // E:548-583: the envs whose candidates are all excluded are smoothed and go on.  Needs the list of those envs on the
    // host (one synchronisation per call: move() is the data-generation API, not a throughput path).
    if (h->smooth_pave) {
        const int N = h->num_envs;
        MG_CUDA(h, cudaMemcpyAsync(h->h_flags, exhausted_dev, (size_t)N, cudaMemcpyDeviceToHost, s));
        MG_CUDA(h, cudaStreamSynchronize(s));
        std::vector<int32_t> list;
        for (int e = 0; e < N; e++)
            if (h->h_flags[e]) list.push_back(e);
        if (!list.empty()) {
            const size_t per = smooth_scratch_bytes(h->P.cap, h->P.ins_cap);
            const size_t want = list.size() < 1024 ? list.size() : 1024;          // envs per launch
            if (h->smooth_slots < want) {
                cudaFree(h->smooth_scratch);
                h->smooth_scratch = nullptr; h->smooth_slots = 0;
                MG_CUDA(h, cudaMalloc((void **)&h->smooth_scratch, per * want));
                h->smooth_slots = want;
            }
            MG_CUDA(h, cudaMemcpyAsync(h->smooth_list, list.data(), sizeof(int32_t) * list.size(), cudaMemcpyHostToDevice, s));
            for (size_t first = 0; first < list.size(); first += h->smooth_slots) {
                const size_t cnt = list.size() - first < h->smooth_slots ? list.size() - first : h->smooth_slots;
                mg_smooth_kernel<<<(unsigned)cnt, 32, h->smem, s>>>(h->P, io, h->smooth_list + first, h->smooth_scratch, per, h->excl_id,
                                                                    h->last_excl_id);
                h->launches++;
            }
            MG_CUDA(h, cudaGetLastError());
            MG_CUDA(h, cudaStreamSynchronize(s));            // `list` is a host object
        }
    }
    return MG_OK;
}

int mg_set_obs_delta(mg_handle h, int enabled) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_set_obs_delta: null handle");
    h->obs_delta = enabled != 0;
    h->obs_bound = nullptr;
    for (auto &b : h->res_bound) b = nullptr;
    return MG_OK;
}

int mg_set_host_delta(mg_handle h, int enabled) { return mg_set_obs_delta(h, enabled); }

int mg_last_host_bytes(mg_handle h, int64_t *h2d, int64_t *d2h) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_last_host_bytes: null handle");
    if (h->acct_valid) {
        // bytes that crossed PCIe in the last mg_step_host: the actions; rewards, flags, element counts for every env;
        // observation rows of the envs that changed (all rows without delta mode or through staging); terminal rows of
        // finished envs (all rows through staging).  The step counters are still those of that step.
        MG_DEVICE(h);
        MG_CUDA(h, cudaMemcpy(h->h_cnt, h->P.counters, CNT_N * sizeof(int32_t), cudaMemcpyDeviceToHost));
        const int64_t N = h->num_envs, row = sizeof(float) * MG_OBS_DIM;
        const int32_t *c = h->h_cnt + CNT_SET * (h->h_cnt[CNT_CUR] & 1);
        int64_t bytes = h->acct_res_delta ? (int64_t)c[CNT_RESBYTES]
                                          : N * (int64_t)(sizeof(double) + 2) + (h->acct_n_elem ? N * (int64_t)sizeof(int32_t) : 0);
        bytes += h->acct_obs_rows_delta ? (int64_t)(c[CNT_OBSERVE] + c[CNT_OBSERVE + 1] + c[CNT_OBSERVE + 2] + c[CNT_OBSERVE + 3] + c[CNT_RESET]) * row : N * row;
        if (h->acct_term_obs) bytes += h->acct_term_obs == 1 ? (int64_t)c[CNT_DONE] * row : N * row;
        h->last_h2d = N * 3 * (int64_t)sizeof(float);
        h->last_d2h = bytes;
        h->acct_valid = false;
    }
    if (h2d) *h2d = h->last_h2d;
    if (d2h) *d2h = h->last_d2h;
    return MG_OK;
}

int mg_sample_actions(mg_handle h, uint64_t seed, uint64_t step_index, float *act_dev, void *stream) {
    if (!h || !act_dev) return fail(h, MG_ERR_ARG, "mg_sample_actions: null pointer");
    MG_DEVICE(h);
    mg_sample_actions_kernel<<<(h->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->num_envs, seed, step_index,
                                                                                         h->P.env_id_offset, act_dev, nullptr);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    return MG_OK;
}

int mg_sample_actions_seq(mg_handle h, uint64_t seed, uint64_t *step_counter_dev, float *act_dev, void *stream) {
    if (!h || !act_dev || !step_counter_dev) return fail(h, MG_ERR_ARG, "mg_sample_actions_seq: null pointer");
    MG_DEVICE(h);
    mg_sample_actions_kernel<<<(h->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        h->num_envs, seed, 0, h->P.env_id_offset, act_dev, reinterpret_cast<unsigned long long *>(step_counter_dev));
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    return MG_OK;
}

int mg_get_state(mg_handle h, int env, mg_state_view *v) {
    if (!h || !v || env < 0 || env >= h->num_envs) return fail(h, MG_ERR_ARG, "mg_get_state: bad argument");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    EnvHot S;
    EnvCold C;
    MG_CUDA(h, cudaMemcpy(&S, h->P.hot + env, sizeof(S), cudaMemcpyDeviceToHost));
    MG_CUDA(h, cudaMemcpy(&C, h->P.cold + env, sizeof(C), cudaMemcpyDeviceToHost));
    v->n = S.n; v->ref_index = S.ref_index; v->n_elements = S.n_elements; v->failed_num = S.failed_num; v->n0 = C.n0;
    v->memo_flags = S.flags;
    v->base_length = S.base_length; v->current_area = S.current_area; v->original_area = C.original_area;
    v->area_min = C.area_min; v->area_crit = C.area_crit;
    const size_t off = (size_t)env * h->P.cap;
    const int n = S.n;
    if (n < 0 || n > h->P.cap) return fail(h, MG_ERR_STATE, "mg_get_state: corrupt env state");
    if (v->xy_host) MG_CUDA(h, cudaMemcpy(v->xy_host, h->P.xy + off, sizeof(double2) * n, cudaMemcpyDeviceToHost));
    if (v->vertex_id_host) MG_CUDA(h, cudaMemcpy(v->vertex_id_host, h->P.vid + off, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    if (v->cand_key_host) MG_CUDA(h, cudaMemcpy(v->cand_key_host, h->P.key + off, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (v->cand_stamp_host) MG_CUDA(h, cudaMemcpy(v->cand_stamp_host, h->P.stamp + off, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    return MG_OK;
}

int mg_get_elements(mg_handle h, int env, int32_t *quads_host, int max_elements, int32_t *n_elements_out,
                    double *vertex_xy_host, int max_vertices, int32_t *n_vertices_out) {
    if (!h || env < 0 || env >= h->num_envs) return fail(h, MG_ERR_ARG, "mg_get_elements: bad argument");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_get_elements: call mg_reset first");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    EnvHot S;
    EnvCold C;
    MG_CUDA(h, cudaMemcpy(&S, h->P.hot + env, sizeof(S), cudaMemcpyDeviceToHost));
    MG_CUDA(h, cudaMemcpy(&C, h->P.cold + env, sizeof(C), cudaMemcpyDeviceToHost));
    bool overflow = false;
    int ne = S.n_elements;
    if (ne > h->P.elem_cap) { ne = h->P.elem_cap; overflow = quads_host != nullptr; }
    if (n_elements_out) *n_elements_out = S.n_elements;
    if (quads_host && max_elements > 0) {
        int c = ne < max_elements ? ne : max_elements;
        MG_CUDA(h, cudaMemcpy(quads_host, h->P.elem + (size_t)env * h->P.elem_cap * 4, sizeof(int32_t) * 4 * c, cudaMemcpyDeviceToHost));
    }
    int nv = C.next_vid;
    if (n_vertices_out) *n_vertices_out = nv;
    if (vertex_xy_host && max_vertices > 0) {
        // original vertices: the domain's template, or (random-polygon mode) the episode's polygon regenerated from
        // its counter; inserted ones from the log
        int n0 = C.n0 < max_vertices ? C.n0 : max_vertices;
        if (!h->P.random_mode) {
            MG_CUDA(h, cudaMemcpy(vertex_xy_host, h->t_xy + (size_t)C.domain * h->P.cap, sizeof(double2) * n0, cudaMemcpyDeviceToHost));
        } else {
            int32_t n_gen = 0;
            const int rc = regen_polygon(h, env, -1, vertex_xy_host, n0, &n_gen, nullptr, nullptr, nullptr, nullptr);
            if (rc != MG_OK) return rc;
            if (n_gen != C.n0) return fail(h, MG_ERR_STATE, "mg_get_elements: regenerated polygon does not match the episode");
        }
        int ni = nv - C.n0;
        if (ni > h->P.ins_cap) { ni = h->P.ins_cap; overflow = true; }
        if (C.n0 + ni > max_vertices) ni = max_vertices - C.n0;
        if (ni > 0)
            MG_CUDA(h, cudaMemcpy(vertex_xy_host + 2 * (size_t)C.n0, h->P.ins_xy + (size_t)env * h->P.ins_cap, sizeof(double2) * ni,
                                  cudaMemcpyDeviceToHost));
    }
    if (overflow)
        return fail(h, MG_ERR_CAPACITY, "mg_get_elements: the episode outgrew the element / inserted-vertex log (the returned "
                                        "prefix and the counts are valid); raise it with mg_set_log_capacity");
    return MG_OK;
}

int mg_debug_polygon(mg_handle h, int env, int episode, double *xy_host, int max_vertices, int32_t *n_out, double *area_out,
                     int32_t *coarse_px_host, int32_t *k_out, double *spacing_out) {
    if (!h || env < 0 || env >= h->num_envs) return fail(h, MG_ERR_ARG, "mg_debug_polygon: bad argument");
    if (!h->P.random_mode) return fail(h, MG_ERR_STATE, "mg_debug_polygon: the handle is not in random-polygon mode");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    return regen_polygon(h, env, episode, xy_host, max_vertices, n_out, area_out, coarse_px_host, k_out, spacing_out);
}

int mg_set_log_capacity(mg_handle h, int max_elements_per_env, int max_inserted_per_env) {
    if (!h || max_elements_per_env < 1 || max_inserted_per_env < 1) return fail(h, MG_ERR_ARG, "mg_set_log_capacity: bad argument");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    Params &P = h->P;
    drop_step_graphs(h);                             // (they were captured with the old logs)
    cudaFree(P.elem); cudaFree(P.ins_xy);
    P.elem = nullptr; P.ins_xy = nullptr;
    h->was_reset = false;            // logs were discarded: the caller resets before stepping again
    P.elem_cap = max_elements_per_env; P.ins_cap = max_inserted_per_env;
    MG_CUDA(h, dalloc(&P.elem, (size_t)h->num_envs * P.elem_cap * 4));
    MG_CUDA(h, dalloc(&P.ins_xy, (size_t)h->num_envs * P.ins_cap));
    return MG_OK;
}

int mg_log_capacity(mg_handle h, int32_t *max_elements_per_env, int32_t *max_inserted_per_env) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_log_capacity: null handle");
    if (max_elements_per_env) *max_elements_per_env = h->P.elem_cap;
    if (max_inserted_per_env) *max_inserted_per_env = h->P.ins_cap;
    return MG_OK;
}

int mg_stats_async(mg_handle h, mg_episode_stats *stats_dev, int reset, void *stream) {
    if (!h || !stats_dev) return fail(h, MG_ERR_ARG, "mg_stats_async: null pointer");
    MG_DEVICE(h);
    mg_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->P.stats, stats_dev, reset);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    note_user_stream(h, (cudaStream_t)stream);
    return MG_OK;
}

int mg_stats(mg_handle h, mg_episode_stats *out, int reset) {
    if (!h || !out) return fail(h, MG_ERR_ARG, "mg_stats: null pointer");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    mg_stats_kernel<<<1, 32>>>(h->P.stats, h->d_stats_out, reset);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    MG_CUDA(h, cudaMemcpy(out, h->d_stats_out, sizeof(*out), cudaMemcpyDeviceToHost));
    return MG_OK;
}

int mg_set_option(mg_handle h, const char *name, int value) {
    if (!h || !name) return fail(h, MG_ERR_ARG, "mg_set_option: null pointer");
    if (std::strcmp(name, "fuse_decide") == 0) { h->fuse_decide = value != 0; return MG_OK; }
    if (std::strcmp(name, "reset_side") == 0) { h->reset_side = value != 0; return MG_OK; }
    if (std::strcmp(name, "pdl") == 0) { h->pdl = value != 0; return MG_OK; }
    if (std::strcmp(name, "host_graph") == 0) { h->host_graph = value != 0; return MG_OK; }
    // not a tuning switch: 0 = mg_move stops where the reference would smooth (env reported done + exhausted)
    if (std::strcmp(name, "smooth_pave") == 0) { h->smooth_pave = value != 0; return MG_OK; }
    // resident one-warp blocks per SM of the item kernels (grid size; default = what fits, see configure_kernels)
    if (std::strcmp(name, "update_blocks") == 0 && value > 0) { h->blocks_update = value; return MG_OK; }
    if (std::strcmp(name, "observe_blocks") == 0 && value > 0) { h->blocks_observe = value; return MG_OK; }
    if (std::strcmp(name, "reset_blocks") == 0 && value > 0) { h->blocks_reset = value; return MG_OK; }
#ifdef MG_TRACE
    // profiling variant: "trace" = capacity arms the item timeline; "trace_dump" writes it to $MESHGEN_TRACE_FILE
    if (std::strcmp(name, "trace") == 0) {
        MG_DEVICE(h);
        MG_CUDA(h, cudaDeviceSynchronize());
        TraceRec *buf = nullptr;
        const unsigned cap = (unsigned)value, zero = 0;
        MG_CUDA(h, cudaMalloc((void **)&buf, sizeof(TraceRec) * (size_t)cap));
        MG_CUDA(h, cudaMemcpyToSymbol(g_trace, &buf, sizeof(buf)));
        MG_CUDA(h, cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap)));
        MG_CUDA(h, cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero)));
        return MG_OK;
    }
    if (std::strcmp(name, "trace_dump") == 0) {
        MG_DEVICE(h);
        MG_CUDA(h, cudaDeviceSynchronize());
        TraceRec *buf = nullptr;
        unsigned cap = 0, n = 0;
        MG_CUDA(h, cudaMemcpyFromSymbol(&buf, g_trace, sizeof(buf)));
        MG_CUDA(h, cudaMemcpyFromSymbol(&cap, g_trace_cap, sizeof(cap)));
        MG_CUDA(h, cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n)));
        if (n > cap) n = cap;
        std::vector<TraceRec> host(n);
        if (n) MG_CUDA(h, cudaMemcpy(host.data(), buf, sizeof(TraceRec) * n, cudaMemcpyDeviceToHost));
        const char *path = std::getenv("MESHGEN_TRACE_FILE");
        FILE *f = std::fopen(path ? path : "trace.bin", "wb");
        if (!f) return fail(h, MG_ERR_ARG, "trace_dump: cannot open the output file");
        std::fwrite(host.data(), sizeof(TraceRec), n, f);
        std::fclose(f);
        const TraceRec *null_buf = nullptr;
        cudaMemcpyToSymbol(g_trace, &null_buf, sizeof(null_buf));
        cudaFree(buf);
        return MG_OK;
    }
#endif
    return fail(h, MG_ERR_ARG, std::string("mg_set_option: unknown option ") + name);
}

int mg_set_kernel_timing(mg_handle h, int enabled) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_set_kernel_timing: null handle");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    if (enabled && !h->ev[0][0])
        for (auto &tr : h->ev)
            for (cudaEvent_t &e : tr) MG_CUDA(h, cudaEventCreate(&e));
    h->timing = enabled != 0;
    h->timing_steps = 0;
    return MG_OK;
}

int mg_kernel_times(mg_handle h, double *ms4, int64_t *steps) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_kernel_times: null handle");
    MG_DEVICE(h);
    MG_CUDA(h, cudaDeviceSynchronize());
    const int64_t n = h->timing_steps < mg_env_s::TIMING_SLOTS ? h->timing_steps : mg_env_s::TIMING_SLOTS;
    double acc[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < 4; k++) {
            float t = 0;
            MG_CUDA(h, cudaEventElapsedTime(&t, h->ev[i][k], h->ev[i][k + 1]));
            acc[k] += t;
        }
    if (ms4)
        for (int k = 0; k < 4; k++) ms4[k] = n ? acc[k] / n : 0;
    if (steps) *steps = n;
    h->timing_steps = 0;
    return MG_OK;
}

}  // extern "C"
