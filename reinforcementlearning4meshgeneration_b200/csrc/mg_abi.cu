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
    DomainScalars *t_sc = nullptr;
    float *t_obs = nullptr;
    mg_episode_stats *d_stats_out = nullptr;
    double2 *sc_tab = nullptr;      // [2][ANGLE_TAB_N] host-libm {sin, cos} of the quantised angles and their halves
    // staging buffers for mg_step_host
    float *d_act = nullptr, *d_obs = nullptr, *d_term_obs = nullptr;
    double *d_rew = nullptr;
    uint8_t *d_term = nullptr, *d_trunc = nullptr;
    int32_t *d_nel = nullptr;
    cudaStream_t host_stream = nullptr;
    // mg_step_host: terminal observations travel compacted (only finished envs)
    int32_t *d_pack_cnt = nullptr, *h_pack_idx = nullptr, *h_pack_cnt = nullptr;
    float *h_pack_obs = nullptr;
    int32_t *h_cnt_all = nullptr;
    float *last_term_obs_host = nullptr;
    // delta mode (mg_set_host_delta): observations / element counts travel only for the envs whose state changed
    bool host_delta = false;
    float *last_obs_host = nullptr;
    int32_t *last_nel_host = nullptr;
    int32_t *h_chg_idx = nullptr, *h_chg_nel = nullptr;
    float *h_chg_obs = nullptr;
    // device-side aliases of the mapped pinned buffers above
    int32_t *m_pack_idx = nullptr, *m_chg_idx = nullptr, *m_chg_nel = nullptr;
    float *m_pack_obs = nullptr, *m_chg_obs = nullptr;
    int64_t last_h2d = 0, last_d2h = 0;
    std::vector<int32_t> prev_done;
    bool ready = false;       // domains or generator configured
    bool was_reset = false;
    int64_t launches = 0;
    int phase_mask = 3;       // profiling aid: bit 0 = phase A launch, bit 1 = phase B+C launch
    int sm_count = 148;
    size_t smem = 0, smem_a = 0;
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

template <class T>
cudaError_t dalloc(T **p, size_t count) {
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, count * sizeof(T));
    return e;
}

int grid_for(int n) { return (n + WPB - 1) / WPB; }

int configure_kernels(mg_handle h) {
    h->smem = smem_bytes(h->P.cap);
    h->smem_a = smem_bytes_a(h->P.cap);
    if (h->smem_a > 48 * 1024)
        MG_CUDA(h, cudaFuncSetAttribute(mg_step_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_a));
    if (h->smem > 48 * 1024) {
        MG_CUDA(h, cudaFuncSetAttribute(mg_step_apply_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        MG_CUDA(h, cudaFuncSetAttribute(mg_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
        MG_CUDA(h, cudaFuncSetAttribute(mg_template_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    }
    // the vertex rings want shared memory, not L1: ask for the largest carve-out so that the number of
    // resident warps is set by registers, not by the driver's default split
    MG_CUDA(h, cudaFuncSetAttribute(mg_step_decide_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    MG_CUDA(h, cudaFuncSetAttribute(mg_step_apply_reset_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
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
    cudaFree(h->t_xy); cudaFree(h->t_key); cudaFree(h->t_stamp); cudaFree(h->t_sc); cudaFree(h->t_obs);
    h->t_xy = nullptr; h->t_key = nullptr; h->t_stamp = nullptr; h->t_sc = nullptr; h->t_obs = nullptr;
}

}  // namespace

extern "C" {

const char *mg_version(void) { return "meshgen_b200 0.1 (sm_100a)"; }

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
    MG_CUDA(h, cudaSetDevice(device));
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    Params &P = h->P;
    P.num_envs = num_envs;
    P.auto_reset = 1;
    P.cap = (max_verts + 1) & ~1;
    const size_t NC = (size_t)num_envs * P.cap;
    P.elem_cap = 8 * P.cap;       // element count per episode scales with the domain's area, not its boundary: the
    P.ins_cap = 8 * P.cap;        // reference's evaluation runs report up to ~5 n0 elements (mg_set_log_capacity to change)
    int rc = MG_OK;
    auto A = [&](cudaError_t er, const char *what) {
        if (er != cudaSuccess && rc == MG_OK) rc = fail(h, MG_ERR_CUDA, std::string("cudaMalloc ") + what + ": " + cudaGetErrorString(er));
    };
    A(dalloc(&P.xy, NC), "xy"); A(dalloc(&P.key, NC), "key"); A(dalloc(&P.stamp, NC), "stamp"); A(dalloc(&P.vid, NC), "vid");
    A(dalloc(&P.st, (size_t)num_envs), "state"); A(dalloc(&P.stats, (size_t)STAT_SLOTS), "stats");
    A(dalloc(&P.obs_cache, (size_t)num_envs * MG_OBS_DIM), "obs");
    A(dalloc(&P.pend, (size_t)num_envs), "pend"); A(dalloc(&P.succ_list, (size_t)num_envs), "succ_list");
    A(dalloc(&P.reset_list, (size_t)num_envs), "reset_list"); A(dalloc(&P.counters, (size_t)CNT_N), "counters");
    A(dalloc(&P.elem, (size_t)num_envs * P.elem_cap * 4), "elem"); A(dalloc(&P.ins_xy, (size_t)num_envs * P.ins_cap), "ins_xy");
    A(dalloc(&h->d_stats_out, 1), "stats_out");
    A(dalloc(&h->sc_tab, (size_t)2 * ANGLE_TAB_N), "angle table");
    A(dalloc(&h->d_act, (size_t)num_envs * 3), "act"); A(dalloc(&h->d_obs, (size_t)num_envs * MG_OBS_DIM), "obs_out");
    A(dalloc(&h->d_term_obs, (size_t)num_envs * MG_OBS_DIM), "term_obs"); A(dalloc(&h->d_rew, (size_t)num_envs), "rew");
    A(dalloc(&h->d_term, (size_t)num_envs), "term"); A(dalloc(&h->d_trunc, (size_t)num_envs), "trunc");
    A(dalloc(&h->d_nel, (size_t)num_envs), "nel");
    A(dalloc(&h->d_pack_cnt, (size_t)2), "pack_cnt");
    // packed rows are written by the pack kernels straight into mapped pinned host memory (coalesced rows over
    // PCIe): no second copy + synchronise round once the counts are known
    A(cudaHostAlloc((void **)&h->h_pack_idx, sizeof(int32_t) * num_envs, cudaHostAllocMapped), "h_pack_idx");
    A(cudaMallocHost((void **)&h->h_pack_cnt, 4 * sizeof(int32_t)), "h_pack_cnt");
    A(cudaMallocHost((void **)&h->h_cnt_all, CNT_N * sizeof(int32_t)), "h_cnt_all");
    A(cudaHostAlloc((void **)&h->h_pack_obs, sizeof(float) * MG_OBS_DIM * num_envs, cudaHostAllocMapped), "h_pack_obs");
    A(cudaHostAlloc((void **)&h->h_chg_idx, sizeof(int32_t) * num_envs, cudaHostAllocMapped), "h_chg_idx");
    A(cudaHostAlloc((void **)&h->h_chg_nel, sizeof(int32_t) * num_envs, cudaHostAllocMapped), "h_chg_nel");
    A(cudaHostAlloc((void **)&h->h_chg_obs, sizeof(float) * MG_OBS_DIM * num_envs, cudaHostAllocMapped), "h_chg_obs");
    if (rc == MG_OK) {
        A(cudaHostGetDevicePointer((void **)&h->m_pack_idx, h->h_pack_idx, 0), "map pack_idx");
        A(cudaHostGetDevicePointer((void **)&h->m_pack_obs, h->h_pack_obs, 0), "map pack_obs");
        A(cudaHostGetDevicePointer((void **)&h->m_chg_idx, h->h_chg_idx, 0), "map chg_idx");
        A(cudaHostGetDevicePointer((void **)&h->m_chg_nel, h->h_chg_nel, 0), "map chg_nel");
        A(cudaHostGetDevicePointer((void **)&h->m_chg_obs, h->h_chg_obs, 0), "map chg_obs");
    }
    if (rc == MG_OK && cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking) != cudaSuccess)
        rc = fail(h, MG_ERR_CUDA, "cudaStreamCreate");
    if (rc == MG_OK) rc = configure_kernels(h);
    if (rc == MG_OK) rc = upload_angle_table(h);
    if (rc != MG_OK) { g_err = h->err; mg_destroy(h); return rc; }
    *out = h;
    return MG_OK;
}

int mg_destroy(mg_handle h) {
    if (!h) return MG_OK;
    cudaSetDevice(h->device);
    Params &P = h->P;
    cudaFree(P.xy); cudaFree(P.key); cudaFree(P.stamp); cudaFree(P.vid); cudaFree(P.st); cudaFree(P.stats);
    cudaFree(P.obs_cache); cudaFree(P.elem); cudaFree(P.ins_xy);
    cudaFree(P.pend); cudaFree(P.succ_list); cudaFree(P.reset_list); cudaFree(P.counters);
    free_templates(h);
    cudaFree(h->sc_tab);
    cudaFree(h->d_stats_out); cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_term_obs); cudaFree(h->d_rew);
    cudaFree(h->d_term); cudaFree(h->d_trunc); cudaFree(h->d_nel);
    cudaFree(h->d_pack_cnt);
    cudaFreeHost(h->h_pack_idx); cudaFreeHost(h->h_pack_cnt); cudaFreeHost(h->h_pack_obs); cudaFreeHost(h->h_cnt_all);
    cudaFreeHost(h->h_chg_idx); cudaFreeHost(h->h_chg_nel); cudaFreeHost(h->h_chg_obs);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
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
    MG_CUDA(h, cudaSetDevice(h->device));
    const size_t N = (size_t)h->num_envs, s = (size_t)slot;
    const int total = h->num_envs * MG_OBS_DIM;
    mg_replay_add_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        h->num_envs, buf_obs + s * N * MG_OBS_DIM, buf_next_obs + s * N * MG_OBS_DIM, buf_act + s * N * MG_ACT_DIM, buf_rew + s * N,
        buf_done + s * N, buf_timeout + s * N, prev_obs, act, new_obs, rew, term, trunc, term_obs);
    h->launches += 1;
    MG_CUDA(h, cudaGetLastError());
    return MG_OK;
}

namespace {
struct Piece { void *ptr; size_t bytes; };
std::vector<Piece> snapshot_pieces(mg_handle h) {
    Params &P = h->P;
    const size_t N = (size_t)h->num_envs, NC = N * P.cap;
    return {
        {P.xy, NC * sizeof(double2)}, {P.key, NC * sizeof(double)}, {P.stamp, NC * sizeof(int32_t)}, {P.vid, NC * sizeof(int32_t)},
        {P.st, N * sizeof(EnvState)}, {P.stats, STAT_SLOTS * sizeof(StatsAcc)}, {P.obs_cache, N * MG_OBS_DIM * sizeof(float)},
        {P.elem, N * P.elem_cap * 4 * sizeof(int32_t)}, {P.ins_xy, N * P.ins_cap * sizeof(double2)},
        {P.counters, CNT_N * sizeof(int)},
    };
}
constexpr size_t SNAP_ALIGN = 256;
size_t snap_round(size_t b) { return (b + SNAP_ALIGN - 1) / SNAP_ALIGN * SNAP_ALIGN; }
}  // namespace

int64_t mg_snapshot_bytes(mg_handle h) {
    if (!h) return 0;
    size_t total = SNAP_ALIGN;                       // header: num_envs, cap, was_reset
    for (const Piece &p : snapshot_pieces(h)) total += snap_round(p.bytes);
    return (int64_t)total;
}

int mg_snapshot_save(mg_handle h, void *blob_dev, void *stream) {
    if (!h || !blob_dev) return fail(h, MG_ERR_ARG, "mg_snapshot_save: null pointer");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_snapshot_save: call mg_reset first");
    MG_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int32_t hdr[4] = {0x4d475331, h->num_envs, h->P.cap, h->P.random_mode};
    MG_CUDA(h, cudaMemcpyAsync(blob_dev, hdr, sizeof(hdr), cudaMemcpyHostToDevice, s));
    MG_CUDA(h, cudaStreamSynchronize(s));            // hdr is a stack object
    char *dst = (char *)blob_dev + SNAP_ALIGN;
    for (const Piece &p : snapshot_pieces(h)) {
        MG_CUDA(h, cudaMemcpyAsync(dst, p.ptr, p.bytes, cudaMemcpyDeviceToDevice, s));
        dst += snap_round(p.bytes);
    }
    return MG_OK;
}

int mg_snapshot_load(mg_handle h, const void *blob_dev, void *stream) {
    if (!h || !blob_dev) return fail(h, MG_ERR_ARG, "mg_snapshot_load: null pointer");
    if (!h->ready) return fail(h, MG_ERR_STATE, "mg_snapshot_load: configure domains or the generator first");
    MG_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int32_t hdr[4] = {0, 0, 0, 0};
    MG_CUDA(h, cudaMemcpyAsync(hdr, blob_dev, sizeof(hdr), cudaMemcpyDeviceToHost, s));
    MG_CUDA(h, cudaStreamSynchronize(s));
    if (hdr[0] != 0x4d475331 || hdr[1] != h->num_envs || hdr[2] != h->P.cap || hdr[3] != h->P.random_mode)
        return fail(h, MG_ERR_ARG, "mg_snapshot_load: blob does not match this handle (num_envs / max_verts / mode)");
    const char *src = (const char *)blob_dev + SNAP_ALIGN;
    for (const Piece &p : snapshot_pieces(h)) {
        MG_CUDA(h, cudaMemcpyAsync(p.ptr, src, p.bytes, cudaMemcpyDeviceToDevice, s));
        src += snap_round(p.bytes);
    }
    h->was_reset = true;
    h->last_obs_host = nullptr;          // host-side delta copies are stale
    h->last_nel_host = nullptr;
    return MG_OK;
}

int mg_set_phase_mask(mg_handle h, int mask) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_set_phase_mask: null handle");
    h->phase_mask = mask & 3;
    // masked steps leave entries in the work lists: start clean on every change of the mask
    MG_CUDA(h, cudaSetDevice(h->device));
    MG_CUDA(h, cudaDeviceSynchronize());
    MG_CUDA(h, cudaMemset(h->P.counters, 0, 4 * sizeof(int)));
    return MG_OK;
}

int mg_num_envs(mg_handle h) { return h ? h->num_envs : 0; }
int mg_max_verts(mg_handle h) { return h ? h->max_verts : 0; }
int64_t mg_launch_count(mg_handle h) { return h ? h->launches : 0; }

int mg_set_domains(mg_handle h, const double *xy_host, const int32_t *offsets_host, int n_domains,
                   const int32_t *env_domain_host, const double *areas_host) {
    if (!h || !xy_host || !offsets_host || !env_domain_host || n_domains <= 0) return fail(h, MG_ERR_ARG, "mg_set_domains: bad argument");
    MG_CUDA(h, cudaSetDevice(h->device));
    Params &P = h->P;
    const int cap = P.cap;
    std::vector<double2> txy((size_t)n_domains * cap, make_double2(0, 0));
    std::vector<DomainScalars> tsc(n_domains);
    for (int d = 0; d < n_domains; d++) {
        int n = offsets_host[d + 1] - offsets_host[d];
        if (n < 3) return fail(h, MG_ERR_ARG, "mg_set_domains: polygon with fewer than 3 vertices");
        if (n > h->max_verts) return fail(h, MG_ERR_CAPACITY, "mg_set_domains: polygon larger than max_verts");
        for (int j = 0; j < n; j++) {
            const double *p = xy_host + 2 * ((size_t)offsets_host[d] + j);
            txy[(size_t)d * cap + j] = make_double2(p[0], p[1]);
        }
        std::memset(&tsc[d], 0, sizeof(DomainScalars));
        tsc[d].n0 = n;
    }
    std::vector<EnvState> st(h->num_envs);
    std::memset(st.data(), 0, sizeof(EnvState) * st.size());
    for (int e = 0; e < h->num_envs; e++) {
        int d = env_domain_host[e];
        if (d < 0 || d >= n_domains) return fail(h, MG_ERR_ARG, "mg_set_domains: env_domain out of range");
        st[e].domain = d;
    }
    free_templates(h);
    MG_CUDA(h, dalloc(&h->t_xy, (size_t)n_domains * cap));
    MG_CUDA(h, dalloc(&h->t_key, (size_t)n_domains * cap));
    MG_CUDA(h, dalloc(&h->t_stamp, (size_t)n_domains * cap));
    MG_CUDA(h, dalloc(&h->t_sc, (size_t)n_domains));
    MG_CUDA(h, dalloc(&h->t_obs, (size_t)n_domains * MG_OBS_DIM));
    MG_CUDA(h, cudaMemcpy(h->t_xy, txy.data(), sizeof(double2) * txy.size(), cudaMemcpyHostToDevice));
    MG_CUDA(h, cudaMemcpy(h->t_sc, tsc.data(), sizeof(DomainScalars) * tsc.size(), cudaMemcpyHostToDevice));
    MG_CUDA(h, cudaMemcpy(P.st, st.data(), sizeof(EnvState) * st.size(), cudaMemcpyHostToDevice));
    double *d_areas = nullptr;
    if (areas_host) {
        MG_CUDA(h, cudaMalloc((void **)&d_areas, sizeof(double) * n_domains));
        MG_CUDA(h, cudaMemcpy(d_areas, areas_host, sizeof(double) * n_domains, cudaMemcpyHostToDevice));
    }
    P.n_domains = n_domains; P.random_mode = 0;
    P.t_xy = h->t_xy; P.t_key = h->t_key; P.t_stamp = h->t_stamp; P.t_sc = h->t_sc; P.t_obs = h->t_obs;
    mg_template_kernel<<<grid_for(n_domains), WPB * 32, h->smem>>>(P, h->t_xy, h->t_key, h->t_stamp, h->t_sc, h->t_obs, d_areas);
    h->launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (d_areas) cudaFree(d_areas);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, MG_ERR_CUDA, std::string("mg_template_kernel: ") + cudaGetErrorString(e));
    h->ready = true; h->was_reset = false;
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
    MG_CUDA(h, cudaSetDevice(h->device));
    MG_CUDA(h, cudaMemset(h->P.st, 0, sizeof(EnvState) * h->num_envs));
    h->P.random_mode = 1; h->P.seed = seed; h->P.gen = c; h->P.env_id_offset = env_id_offset;
    h->ready = true; h->was_reset = false;
    return MG_OK;
}

int mg_reset(mg_handle h, const uint8_t *mask_dev, float *obs_dev, void *stream) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_reset: null handle");
    if (!h->ready) return fail(h, MG_ERR_STATE, "mg_reset: call mg_set_domains or mg_set_random first");
    if (mask_dev && !h->was_reset) return fail(h, MG_ERR_STATE, "mg_reset: the first reset must cover all envs (mask = NULL)");
    MG_CUDA(h, cudaSetDevice(h->device));
    mg_reset_kernel<<<grid_for(h->num_envs), WPB * 32, h->smem, (cudaStream_t)stream>>>(h->P, mask_dev, obs_dev);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    h->was_reset = true;
    h->last_obs_host = nullptr;          // host-side delta copies are stale after a reset
    h->last_nel_host = nullptr;
    return MG_OK;
}

int mg_step(mg_handle h, const float *act_dev, float *obs_dev, double *rew_dev, uint8_t *term_dev, uint8_t *trunc_dev,
            float *term_obs_dev, int32_t *n_elem_dev, void *stream) {
    if (!h || !act_dev || !obs_dev || !rew_dev || !term_dev || !trunc_dev) return fail(h, MG_ERR_ARG, "mg_step: null pointer");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_step: call mg_reset first");
    MG_CUDA(h, cudaSetDevice(h->device));
    StepIO io;
    io.act = act_dev; io.obs_out = obs_dev; io.rew_out = rew_dev; io.term_out = term_dev; io.trunc_out = trunc_dev;
    io.term_obs_out = term_obs_dev; io.n_elem_out = n_elem_dev;
    cudaStream_t s = (cudaStream_t)stream;
    const int full = grid_for(h->num_envs);
    const int gb = full < h->sm_count * (32 / WPB) ? full : h->sm_count * (32 / WPB);     // 32 item slots per SM
    const int gc = full < h->sm_count * (8 / WPB) ? full : h->sm_count * (8 / WPB);
    if (h->phase_mask & 1) mg_step_decide_kernel<<<(h->num_envs + WPB_A - 1) / WPB_A, WPB_A * 32, h->smem_a, s>>>(h->P, io);
    if (h->phase_mask & 2) mg_step_apply_reset_kernel<<<gb + gc, WPB * 32, h->smem, s>>>(h->P, io, gb);
    h->launches += 2;
    MG_CUDA(h, cudaGetLastError());
    return MG_OK;
}

int mg_step_host(mg_handle h, const float *act_host, float *obs_host, double *rew_host, uint8_t *term_host,
                 uint8_t *trunc_host, float *term_obs_host, int32_t *n_elem_host) {
    if (!h || !act_host || !obs_host || !rew_host || !term_host || !trunc_host) return fail(h, MG_ERR_ARG, "mg_step_host: null pointer");
    if (!h->was_reset) return fail(h, MG_ERR_STATE, "mg_step_host: call mg_reset first");
    MG_CUDA(h, cudaSetDevice(h->device));
    const size_t N = h->num_envs;
    cudaStream_t s = h->host_stream;
    MG_CUDA(h, cudaMemcpyAsync(h->d_act, act_host, N * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = mg_step(h, h->d_act, h->d_obs, h->d_rew, h->d_term, h->d_trunc, h->d_term_obs, h->d_nel, s);
    if (rc != MG_OK) return rc;
    int64_t d2h = 0;
    // observations (and element counts) only change for envs that created an element or were reset: in delta
    // mode, with the caller's buffers unchanged since the previous call, only those rows cross PCIe
    const bool delta_obs = h->host_delta && h->last_obs_host == obs_host;
    // (element counts are not delta-coded: a reset env reports the finished episode's count in the reset
    // step and 0 from the next step on, whatever that step does)
    const bool delta_nel = false;
    // pinned caller buffers: changed rows go straight into them (no staging, no host-side scatter)
    float *obs_alias = delta_obs ? pinned_alias(obs_host) : nullptr;
    float *tobs_alias = pinned_alias(term_obs_host);
    const bool direct_obs = obs_alias != nullptr, direct_tobs = tobs_alias != nullptr;
    MG_CUDA(h, cudaMemsetAsync(h->d_pack_cnt, 0, 2 * sizeof(int32_t), s));
    if (direct_obs || direct_tobs) {
        mg_scatter_rows_host_kernel<<<h->sm_count * 2, 256, 0, s>>>(h->P, h->d_obs, direct_obs ? obs_alias : nullptr, h->d_term,
                                                                   h->d_trunc, h->d_term_obs, direct_tobs ? tobs_alias : nullptr,
                                                                   h->d_pack_cnt);
        h->launches++;
    }
    if (term_obs_host && !direct_tobs) {
        mg_pack_terminal_kernel<<<(h->num_envs + 255) / 256, 256, 0, s>>>(h->num_envs, h->d_term, h->d_trunc, h->d_term_obs,
                                                                           h->m_pack_idx, h->m_pack_obs, h->d_pack_cnt);
        h->launches++;
    }
    if (delta_obs && !direct_obs) {
        mg_pack_changed_kernel<<<h->sm_count * 2, 256, 0, s>>>(h->P, h->d_obs, h->d_nel, h->m_chg_idx, h->m_chg_obs,
                                                              h->m_chg_nel, h->d_pack_cnt + 1);
        h->launches++;
    } else if (!delta_obs) {
        MG_CUDA(h, cudaMemcpyAsync(obs_host, h->d_obs, N * MG_OBS_DIM * sizeof(float), cudaMemcpyDeviceToHost, s));
        d2h += N * MG_OBS_DIM * sizeof(float);
    }
    if (direct_obs || direct_tobs) {      // row counts for the byte accounting (both counter sets + the current set)
        MG_CUDA(h, cudaMemcpyAsync(h->h_cnt_all, h->P.counters, CNT_N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    MG_CUDA(h, cudaMemcpyAsync(h->h_pack_cnt, h->d_pack_cnt, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    MG_CUDA(h, cudaMemcpyAsync(rew_host, h->d_rew, N * sizeof(double), cudaMemcpyDeviceToHost, s));
    MG_CUDA(h, cudaMemcpyAsync(term_host, h->d_term, N, cudaMemcpyDeviceToHost, s));
    MG_CUDA(h, cudaMemcpyAsync(trunc_host, h->d_trunc, N, cudaMemcpyDeviceToHost, s));
    d2h += N * (sizeof(double) + 2) + 2 * sizeof(int32_t);
    if (n_elem_host && !delta_nel) {
        MG_CUDA(h, cudaMemcpyAsync(n_elem_host, h->d_nel, N * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        d2h += N * sizeof(int32_t);
    }
    MG_CUDA(h, cudaStreamSynchronize(s));
    const int c_term = (term_obs_host && !direct_tobs) ? h->h_pack_cnt[0] : 0, c_chg = (delta_obs && !direct_obs) ? h->h_pack_cnt[1] : 0;
    d2h += (int64_t)c_term * (4 + 4 * MG_OBS_DIM) + (int64_t)c_chg * (4 + 4 * MG_OBS_DIM + 4);    // written by the pack kernels
    if (direct_obs) {
        const int cur = h->h_cnt_all[CNT_CUR] & 1;
        d2h += (int64_t)(h->h_cnt_all[2 * cur] + h->h_cnt_all[2 * cur + 1]) * 4 * MG_OBS_DIM;
    }
    if (direct_tobs) d2h += (int64_t)h->h_pack_cnt[0] * 4 * MG_OBS_DIM;
    for (int i = 0; i < c_chg; i++) {
        const size_t e = (size_t)h->h_chg_idx[i];
        std::memcpy(obs_host + e * MG_OBS_DIM, h->h_chg_obs + (size_t)i * MG_OBS_DIM, sizeof(float) * MG_OBS_DIM);
        if (delta_nel) n_elem_host[e] = h->h_chg_nel[i];
    }
    if (term_obs_host && !direct_tobs) {
        // terminal observations are only defined where done: those rows travel compacted instead of N*72 bytes
        if (h->last_term_obs_host != term_obs_host) {
            std::memset(term_obs_host, 0, sizeof(float) * MG_OBS_DIM * N);
            h->last_term_obs_host = term_obs_host;
        } else {
            for (int32_t e : h->prev_done) std::memset(term_obs_host + (size_t)e * MG_OBS_DIM, 0, sizeof(float) * MG_OBS_DIM);
        }
        h->prev_done.assign(h->h_pack_idx, h->h_pack_idx + c_term);
        for (int i = 0; i < c_term; i++)
            std::memcpy(term_obs_host + (size_t)h->h_pack_idx[i] * MG_OBS_DIM, h->h_pack_obs + (size_t)i * MG_OBS_DIM,
                        sizeof(float) * MG_OBS_DIM);
    }
    h->last_obs_host = obs_host;
    h->last_nel_host = n_elem_host;
    h->last_h2d = (int64_t)(N * 3 * sizeof(float));
    h->last_d2h = d2h;
    return MG_OK;
}

int mg_set_host_delta(mg_handle h, int enabled) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_set_host_delta: null handle");
    h->host_delta = enabled != 0;
    h->last_obs_host = nullptr;
    h->last_nel_host = nullptr;
    return MG_OK;
}

int mg_last_host_bytes(mg_handle h, int64_t *h2d, int64_t *d2h) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_last_host_bytes: null handle");
    if (h2d) *h2d = h->last_h2d;
    if (d2h) *d2h = h->last_d2h;
    return MG_OK;
}

int mg_sample_actions(mg_handle h, uint64_t seed, uint64_t step_index, float *act_dev, void *stream) {
    if (!h || !act_dev) return fail(h, MG_ERR_ARG, "mg_sample_actions: null pointer");
    MG_CUDA(h, cudaSetDevice(h->device));
    mg_sample_actions_kernel<<<(h->num_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->num_envs, seed, step_index,
                                                                                         h->P.env_id_offset, act_dev);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    return MG_OK;
}

int mg_get_state(mg_handle h, int env, mg_state_view *v) {
    if (!h || !v || env < 0 || env >= h->num_envs) return fail(h, MG_ERR_ARG, "mg_get_state: bad argument");
    MG_CUDA(h, cudaSetDevice(h->device));
    MG_CUDA(h, cudaDeviceSynchronize());
    EnvState S;
    MG_CUDA(h, cudaMemcpy(&S, h->P.st + env, sizeof(S), cudaMemcpyDeviceToHost));
    v->n = S.n; v->ref_index = S.ref_index; v->n_elements = S.n_elements; v->failed_num = S.failed_num; v->n0 = S.n0;
    v->base_length = S.base_length; v->current_area = S.current_area; v->original_area = S.original_area;
    v->area_min = S.area_min; v->area_crit = S.area_crit;
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
    MG_CUDA(h, cudaSetDevice(h->device));
    MG_CUDA(h, cudaDeviceSynchronize());
    EnvState S;
    MG_CUDA(h, cudaMemcpy(&S, h->P.st + env, sizeof(S), cudaMemcpyDeviceToHost));
    int ne = S.n_elements < h->P.elem_cap ? S.n_elements : h->P.elem_cap;
    if (n_elements_out) *n_elements_out = S.n_elements;
    if (quads_host && max_elements > 0) {
        int c = ne < max_elements ? ne : max_elements;
        MG_CUDA(h, cudaMemcpy(quads_host, h->P.elem + (size_t)env * h->P.elem_cap * 4, sizeof(int32_t) * 4 * c, cudaMemcpyDeviceToHost));
    }
    int nv = S.next_vid;
    if (n_vertices_out) *n_vertices_out = nv;
    if (vertex_xy_host && max_vertices > 0) {
        // original vertices come from the template (domain mode); inserted ones from the log
        int n0 = S.n0 < max_vertices ? S.n0 : max_vertices;
        if (!h->P.random_mode)
            MG_CUDA(h, cudaMemcpy(vertex_xy_host, h->t_xy + (size_t)S.domain * h->P.cap, sizeof(double2) * n0, cudaMemcpyDeviceToHost));
        else std::memset(vertex_xy_host, 0, sizeof(double) * 2 * n0);
        int ni = nv - S.n0;
        if (ni > h->P.ins_cap) ni = h->P.ins_cap;
        if (S.n0 + ni > max_vertices) ni = max_vertices - S.n0;
        if (ni > 0)
            MG_CUDA(h, cudaMemcpy(vertex_xy_host + 2 * (size_t)S.n0, h->P.ins_xy + (size_t)env * h->P.ins_cap, sizeof(double2) * ni,
                                  cudaMemcpyDeviceToHost));
    }
    return MG_OK;
}

int mg_set_log_capacity(mg_handle h, int max_elements_per_env, int max_inserted_per_env) {
    if (!h || max_elements_per_env < 1 || max_inserted_per_env < 1) return fail(h, MG_ERR_ARG, "mg_set_log_capacity: bad argument");
    MG_CUDA(h, cudaSetDevice(h->device));
    MG_CUDA(h, cudaDeviceSynchronize());
    Params &P = h->P;
    cudaFree(P.elem); cudaFree(P.ins_xy);
    P.elem = nullptr; P.ins_xy = nullptr;
    P.elem_cap = max_elements_per_env; P.ins_cap = max_inserted_per_env;
    MG_CUDA(h, dalloc(&P.elem, (size_t)h->num_envs * P.elem_cap * 4));
    MG_CUDA(h, dalloc(&P.ins_xy, (size_t)h->num_envs * P.ins_cap));
    h->was_reset = false;            // logs were discarded: the caller resets before stepping again
    return MG_OK;
}

int mg_log_capacity(mg_handle h, int32_t *max_elements_per_env, int32_t *max_inserted_per_env) {
    if (!h) return fail(h, MG_ERR_ARG, "mg_log_capacity: null handle");
    if (max_elements_per_env) *max_elements_per_env = h->P.elem_cap;
    if (max_inserted_per_env) *max_inserted_per_env = h->P.ins_cap;
    return MG_OK;
}

int mg_stats(mg_handle h, mg_episode_stats *out, int reset) {
    if (!h || !out) return fail(h, MG_ERR_ARG, "mg_stats: null pointer");
    MG_CUDA(h, cudaSetDevice(h->device));
    MG_CUDA(h, cudaDeviceSynchronize());
    mg_stats_kernel<<<1, 32>>>(h->P.stats, h->d_stats_out, reset);
    h->launches++;
    MG_CUDA(h, cudaGetLastError());
    MG_CUDA(h, cudaMemcpy(out, h->d_stats_out, sizeof(*out), cudaMemcpyDeviceToHost));
    return MG_OK;
}

}  // extern "C"
