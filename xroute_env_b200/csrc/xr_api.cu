// xr_api.cu -- C ABI (include/xroute_b200.h) and host-side step state machine of the
// B200-native XRoute environment hot path.  No torch types; CUDA runtime only.
//
// Host mirror of reference semantics:
//   Game.reset  baseline/baseline_utils.py:441-481   -> xr_reset
//   Game.step   baseline/baseline_utils.py:392-439   -> xr_step / xr_step_results
//   legal set   baseline_utils.py:438,472            -> host mirror h_routed/h_has_ap
#include "../../include/xroute_b200.h"
#include "dlpack_abi.h"
#include "xr_common.cuh"
#include "xr_kernels_env.cuh"
#include "xr_kernels_maze.cuh"
#include "xr_kernels_win.cuh"
#include "xr_kernels_win2.cuh"
#include "xr_frontier.h"

#include <algorithm>
#include <dlfcn.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static std::string g_create_error;

#define XR_NG 3   // post-route groups
#define XR_NB 9   // launch buckets of the window kernels: band layout with 1, 2, 4, 8, 16 CTAs, then dual cyclic layout with 2, 4, 8, 16
struct ProfEvent { int cls; cudaEvent_t a, b; cudaEvent_t step; int grp; };

struct XrEnv {
    XrConfig cfg;
    Geo g;
    Dev d;
    int device = 0;
    std::string err;
    std::vector<void *> allocs;
    // host copies of geometry
    std::vector<int32_t> xc, yc;
    // host mirrors
    std::vector<uint8_t> h_routed, h_has_ap, h_done, h_loaded, h_reset;
    std::vector<uint16_t> h_npins;      // distinct pins per net
    std::vector<int32_t> h_naps;        // access points per net
    std::vector<int32_t> h_nrem;
    int32_t *p_act = nullptr;           // pinned [N][2]
    int32_t *p_flags = nullptr;         // pinned [2]
    unsigned char *p_res = nullptr;     // pinned mirror of the step results: delta int32[N][3], cum int64[N][6], done u8[N]
    bool res_on_host = false;           // p_res holds the results of the last step (filled with the step's own read-back)
    size_t rb_bytes = 0;                // size of the device read-back block (cum | delta | flags | done)
    int32_t *p_ids = nullptr;           // pinned [N]
    uint8_t *p_full = nullptr;          // pinned [N] reset: full observation build needed
    std::vector<uint8_t> h_clean;       // observation buffer of the env satisfies the incremental invariant
    int obs_mode = 0;
    int32_t *d_ids = nullptr;
    int pumps_per_sync = 4;
    // window-resident route kernel
    int win_margin = 14, min_cluster = 0, smem_cap = 0, n_sm = 148;
    int win_fit_cap = 0;                // bytes of shared memory a window may take per CTA (smem_cap; XR_WIN_FIT_CAP lowers it: tests)
    std::vector<int32_t> h_netwin;      // [N][max_nets+1][2]  WX, WY  (0 = no window)
    int32_t *p_lists = nullptr;         // pinned [3 + XR_NB*XR_NG][N]: mode, group, frontier list, env lists of the XR_NG x XR_NB cluster buckets
    int32_t *d_lists = nullptr;         // device [1 + XR_NB*XR_NG][N]
    cudaStream_t gs[XR_NG] = {nullptr, nullptr, nullptr};   // one stream per post-route group
    cudaEvent_t ev_fork = nullptr, ev_join[XR_NG] = {nullptr, nullptr, nullptr};
    cudaStream_t s_glob = nullptr;      // the full-grid sweeps of a step run here, beside the frontier / window kernels
    cudaEvent_t ev_glob = nullptr;      // ... from the moment the prologue of their environments is done
    int grp_pins[XR_NG] = {0, 4, 8};    // group g = nets with at least grp_pins[g] pins (light / medium / heavy)
    int heavy_cluster = 8;              // minimum cluster size of the heaviest group (0 = same as the others)
    int dual_pins = 8;                  // nets with at least this many pins use the dual cyclic layout (0 = never)
    int dual_minc = 8;                  // ... on clusters of at least this many CTAs
    long long n_win_nets = 0, n_global_nets = 0, n_frontier_nets = 0;
    // frontier engine (default): one CTA per net, goal-directed search on the epoch-tagged global field
    int engine = 0;                     // 0 = frontier, 1 = window kernels + full-grid sweeps (round-1 engines)
    FrParams fr_big = {}, fr_small = {};// list capacities for "one CTA per SM" and "several CTAs per SM" launches
    int fr_threads_big = FR_T, fr_threads_small = 512;
    int hybrid_area = 4000, hybrid_pins = 5;   // hybrid: nets of at most hybrid_pins pins whose access-point box covers at least hybrid_area cells take the sweep kernels (0 = never)
    std::vector<int32_t> h_area;        // [N][max_nets+1] cells of the net's access-point bounding box (x by y)
    int guide_cap = 0;                  // guide boxes per environment the device table holds (grown on demand)
    int metrics_mode = 0;               // 0 = congestion counts maintained by the commits, 1 = full scan (k_metrics) every step
    // counters
    long long n_launch = 0, n_sync = 0;
    // profiling
    bool prof = false;
    std::vector<ProfEvent> prof_pending;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[XR_K_COUNT] = {0};
    cudaEvent_t cur_step_ev = nullptr;  // start of the current step (profiling timeline)
    int cur_grp = -1;
    std::vector<cudaEvent_t> step_evs;
    double tl_sum[XR_NG][3] = {};   // per group: route start / route end / obs end offsets (ms)
    long long tl_n[XR_NG][3] = {};
    long long prof_n[XR_K_COUNT] = {0};
    // a step enqueued by xr_step_async and not yet completed by xr_step_wait
    struct { bool active = false, any_route = false, any_global = false, any_win = false; int maxn = 0; cudaStream_t st = nullptr; } pend;
    cudaEvent_t ev_done = nullptr;
    std::atomic<int> refs{1};           // handle + outstanding DLPack tensors
};

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            env->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
            return XR_E_CUDA;                                                             \
        }                                                                                 \
    } while (0)

static int fail(XrEnv *env, int code, const std::string &msg) {
    if (env) env->err = msg; else g_create_error = msg;
    return code;
}

template <typename T>
static cudaError_t dalloc(XrEnv *env, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    e = cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T));
    env->allocs.push_back(q);
    *p = reinterpret_cast<T *>(q);
    return e;
}

// -------------------------------------------------------------- launch helpers
struct Launch {
    XrEnv *env; int cls; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    Launch(XrEnv *e, int c, cudaStream_t s) : env(e), cls(c), st(s) {
        env->n_launch++;
        if (env->prof) {
            a = get(); b = get();
            cudaEventRecord(a, st);
        }
    }
    cudaEvent_t get() {
        if (!env->ev_pool.empty()) { cudaEvent_t e = env->ev_pool.back(); env->ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    ~Launch() {
        if (env->prof) { cudaEventRecord(b, st); env->prof_pending.push_back({cls, a, b, env->cur_step_ev, env->cur_grp}); }
    }
};

static void prof_collect(XrEnv *env) {
    for (auto &p : env->prof_pending) {
        float ms = 0.f;
        cudaEventSynchronize(p.b);
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { env->prof_ms[p.cls] += ms; env->prof_n[p.cls]++; }
        if (p.step && p.grp >= 0 && p.grp < XR_NG && (p.cls == XR_K_ROUTE_WIN || p.cls == XR_K_ROUTE_FRONTIER || p.cls == XR_K_OBS)) {
            float o0 = 0.f, o1 = 0.f;
            if (cudaEventElapsedTime(&o0, p.step, p.a) == cudaSuccess && cudaEventElapsedTime(&o1, p.step, p.b) == cudaSuccess) {
                if (p.cls != XR_K_OBS) {
                    env->tl_sum[p.grp][0] += o0; env->tl_n[p.grp][0]++;
                    env->tl_sum[p.grp][1] += o1; env->tl_n[p.grp][1]++;
                } else { env->tl_sum[p.grp][2] += o1; env->tl_n[p.grp][2]++; }
            }
        }
        env->ev_pool.push_back(p.a); env->ev_pool.push_back(p.b);
    }
    env->prof_pending.clear();
    for (auto e : env->step_evs) env->ev_pool.push_back(e);
    env->step_evs.clear();
    env->cur_step_ev = nullptr;
}

// ----------------------------------------------------------------- create/destroy
extern "C" int xr_version(void) { return XR_VERSION; }
extern "C" int xr_step_wait(XrEnv *env);

extern "C" const char *xr_last_error(const XrEnv *env) {
    return env ? env->err.c_str() : g_create_error.c_str();
}

static void xr_free(XrEnv *env) {
    cudaSetDevice(env->device);
    for (void *p : env->allocs) cudaFree(p);
    if (env->p_act) cudaFreeHost(env->p_act);
    if (env->p_flags) cudaFreeHost(env->p_flags);
    if (env->p_res) cudaFreeHost(env->p_res);
    if (env->p_ids) cudaFreeHost(env->p_ids);
    if (env->p_full) cudaFreeHost(env->p_full);

    for (int k = 0; k < XR_NG; k++) { if (env->gs[k]) cudaStreamDestroy(env->gs[k]); if (env->ev_join[k]) cudaEventDestroy(env->ev_join[k]); }
    if (env->ev_fork) cudaEventDestroy(env->ev_fork);
    if (env->s_glob) cudaStreamDestroy(env->s_glob);
    if (env->ev_glob) cudaEventDestroy(env->ev_glob);
    if (env->ev_done) cudaEventDestroy(env->ev_done);
    for (auto &p : env->prof_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : env->ev_pool) cudaEventDestroy(e);
    delete env;
}
static void xr_unref(XrEnv *env) {
    if (env->refs.fetch_sub(1) == 1) xr_free(env);
}

extern "C" void xr_destroy(XrEnv *env) {
    if (!env) return;
    cudaSetDevice(env->device);
    cudaDeviceSynchronize();
    xr_unref(env);
}

extern "C" int xr_create(const XrConfig *cfg, XrEnv **out) {
    if (!cfg || !out) return fail(nullptr, XR_E_INVALID, "null argument");
    if (cfg->n_envs < 1 || cfg->X < 2 || cfg->Y < 2 || cfg->Z < 1 || cfg->Z > XR_MAX_LAYERS ||
        cfg->X > 1024 || cfg->Y > 1024 || cfg->max_nets < 1 || cfg->max_nets > 65535 || cfg->max_aps < 1)
        return fail(nullptr, XR_E_INVALID, "grid/net limits: 2<=X,Y<=1024, 1<=Z<=16, 1<=max_nets<=65535");
    if (!cfg->x_coords || !cfg->y_coords || !cfg->layer_dir || !cfg->layer_pitch || !cfg->layer_min_width)
        return fail(nullptr, XR_E_INVALID, "geometry arrays missing");
    if ((long long)cfg->X * cfg->Y * cfg->Z >= (1ll << 30))
        return fail(nullptr, XR_E_INVALID, "grid too large");
    if (cfg->via_cost < 1 || cfg->grid_cost < 0 || cfg->drc_cost < 0 || cfg->fixed_shape_cost < 0 || cfg->block_cost < 0 ||
        1 + cfg->grid_cost + cfg->drc_cost + cfg->fixed_shape_cost > 255)
        return fail(nullptr, XR_E_INVALID, "cost constants out of range (1 + grid + drc + fixed must be <= 255)");
    cudaError_t ce = cudaSetDevice(cfg->device);
    if (ce != cudaSuccess) return fail(nullptr, XR_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
    XrEnv *env = new XrEnv();
    env->cfg = *cfg;
    env->device = cfg->device;
    Geo &g = env->g;
    memset(&g, 0, sizeof(g));
    g.N = cfg->n_envs; g.X = cfg->X; g.Y = cfg->Y; g.Z = cfg->Z;
    g.Xp = (cfg->X + 31) / 32 * 32;
    g.cells = cfg->X * cfg->Y * cfg->Z;
    g.cells_p = cfg->Z * cfg->Y * g.Xp;
    g.cells_o = (g.cells + 15) / 16 * 16;
    g.max_nets = cfg->max_nets; g.max_aps = cfg->max_aps;
    g.obs_max_nets = (cfg->obs_max_nets < 0 || cfg->obs_max_nets > cfg->max_nets) ? cfg->max_nets : cfg->obs_max_nets;
    g.path_cap = cfg->path_capacity > 0 ? cfg->path_capacity : 16 * (cfg->X + cfg->Y + cfg->Z) + 1024;
    g.conn_cap = 256;
    const long long maxc = 2ll + 7ll * g.obs_max_nets;
    g.obs_stride = (maxc * g.cells + 63) / 64 * 64;
    env->pumps_per_sync = cfg->pumps_per_sync > 0 ? cfg->pumps_per_sync : 4;
    env->xc.assign(cfg->x_coords, cfg->x_coords + cfg->X);
    env->yc.assign(cfg->y_coords, cfg->y_coords + cfg->Y);
    for (int i = 1; i < cfg->X; i++) if (env->xc[i] <= env->xc[i - 1]) { delete env; return fail(nullptr, XR_E_INVALID, "x_coords not increasing"); }
    for (int i = 1; i < cfg->Y; i++) if (env->yc[i] <= env->yc[i - 1]) { delete env; return fail(nullptr, XR_E_INVALID, "y_coords not increasing"); }
    g.uniform_x = 1; g.dx = env->xc[1] - env->xc[0];
    for (int i = 1; i < cfg->X; i++) if (env->xc[i] - env->xc[i - 1] != g.dx) g.uniform_x = 0;
    g.uniform_y = 1; g.dy = env->yc[1] - env->yc[0];
    for (int i = 1; i < cfg->Y; i++) if (env->yc[i] - env->yc[i - 1] != g.dy) g.uniform_y = 0;
    for (int z = 0; z < cfg->Z; z++) {
        const int horiz = cfg->layer_dir[z] == 0;
        for (int f = 0; f < 4; f++) {
            const uint32_t base = 1u + (uint32_t)cfg->drc_cost * (f & 1) + (uint32_t)cfg->fixed_shape_cost * ((f >> 1) & 1);
            g.multX[z][f] = base + (horiz ? 0u : (uint32_t)cfg->grid_cost);
            g.multY[z][f] = base + (horiz ? (uint32_t)cfg->grid_cost : 0u);
            g.multV[f] = base;
        }
        g.pen[z] = (uint32_t)cfg->block_cost * (uint32_t)cfg->layer_min_width[z] * 20u;
        g.vlen[z] = (z + 1 < cfg->Z) ? (uint32_t)cfg->via_cost * (uint32_t)cfg->layer_pitch[z + 1] : 0u;
    }
    const size_t N = g.N;
#define DA(ptr, n)                                                                        \
    do {                                                                                  \
        ce = dalloc(env, &(ptr), (size_t)(n));                                            \
        if (ce != cudaSuccess) {                                                          \
            std::string m = std::string("cudaMalloc " #ptr ": ") + cudaGetErrorString(ce); \
            xr_free(env);                                                                 \
            return fail(nullptr, XR_E_CUDA, m);                                           \
        }                                                                                 \
    } while (0)
    Dev &d = env->d;
    int32_t *dxc, *dyc;
    DA(dxc, g.X); DA(dyc, g.Y);
    cudaMemcpy(dxc, env->xc.data(), sizeof(int32_t) * g.X, cudaMemcpyHostToDevice);
    cudaMemcpy(dyc, env->yc.data(), sizeof(int32_t) * g.Y, cudaMemcpyHostToDevice);
    g.xc = dxc; g.yc = dyc;
    DA(d.cellinfo, N * g.cells_p); DA(d.apnet, N * g.cells_p);
    DA(d.ap_cellp, N * g.max_aps); DA(d.ap_obsoff, N * g.max_aps); DA(d.ap_pin, N * g.max_aps);
    DA(d.ap_adj, N * g.max_aps); DA(d.net_start, N * (g.max_nets + 2)); DA(d.net_srcpin, N * (g.max_nets + 1));
    DA(d.obst_obs, N * g.cells_o); DA(d.routed, N * (g.max_nets + 1)); DA(d.legal, N * (g.max_nets + 1));
    DA(d.rank_net, N * g.max_nets); DA(d.n_remaining, N);
    DA(d.dist, N * g.cells_p); DA(d.cflag, N * g.cells_p);
    DA(d.phase, N); DA(d.changed, N); DA(d.reinit, N); DA(d.first, N);
    {   // everything a step uploads lives in one block (one H2D copy): act [N][2] | mode [N] | grp [N] | env lists
        int32_t *up;
        DA(up, N * (5 + XR_NB * XR_NG));
        d.act = up; d.mode = up + 2 * N; d.grp = up + 3 * N; env->d_lists = up + 4 * N;
    }
    {   // ... and everything a step reads back in another (one D2H copy): cum i64 [N][6] | delta i32 [N][3] | flags i32 [4] | done u8 [N]
        int64_t *rb;
        DA(rb, N * XR_M_COUNT + (N * 3 + 4 + 1) / 2 + (N + 7) / 8 + 1);
        d.cum = reinterpret_cast<long long *>(rb);
        d.delta = reinterpret_cast<int32_t *>(rb + N * XR_M_COUNT);
        d.flags = d.delta + 3 * N;
        d.done = reinterpret_cast<uint8_t *>(d.flags + 4);
        env->rb_bytes = sizeof(int64_t) * N * XR_M_COUNT + sizeof(int32_t) * (3 * N + 4) + N;
    }
    DA(d.ap_conn, N * g.max_aps);
    DA(d.g_rowd, N * g.Y); DA(d.g_rowf, N * g.Y); DA(d.g_slabd, N * g.Z * (g.Xp / 32)); DA(d.g_slabf, N * g.Z * (g.Xp / 32));
    DA(d.g_cap, N); DA(d.g_gmin, N); DA(d.g_all, N);
    DA(d.msum, N * 4); DA(d.wlvia, N * 2);
    DA(d.reward, N); DA(d.envstat, N * 8); DA(d.stats, XR_STATS_COUNT); DA(d.obs_do, N); DA(d.obs_full, N);
    DA(d.path, N * g.path_cap); DA(d.path_n, N); DA(d.conn_off, N * (g.conn_cap + 1));
    DA(d.conn_cost, N * g.conn_cap); DA(d.conn_n, N);
    DA(env->d_ids, N);
    {   // frontier engine: epoch-tagged distance field (all ones = older than any epoch), list spill space
        const int cp = g.cells_p;
        env->fr_big.cap_s = 6144; env->fr_big.cap_e = 4096;
        env->fr_small.cap_s = 2048; env->fr_small.cap_e = 1024;
        env->fr_big.delta = env->fr_small.delta = 400u; env->fr_big.ray = env->fr_small.ray = FR_RAY;
        if (const char *e = getenv("XR_FR_DELTA")) env->fr_big.delta = env->fr_small.delta = (uint32_t)std::max(0, atoi(e));
        env->fr_big.dmax = env->fr_small.dmax = 4;
        if (const char *e = getenv("XR_FR_DMAX")) env->fr_big.dmax = env->fr_small.dmax = std::max(1, atoi(e));
        // the wide variant of the kernel (far list from 4096 open entries on, band = one full-width bucket) for grids of more
        // than 2 M cells: their wide nets hold 10^5 open entries; smaller grids keep the shorter round of the plain variant
        int park = cp > (2 << 20) ? 4096 : 0, band = 1;
        if (const char *e = getenv("XR_FR_PARK")) park = std::max(0, atoi(e));
        if (const char *e = getenv("XR_FR_BAND")) band = std::max(1, atoi(e));
        env->fr_big.park_min = env->fr_small.park_min = park;
        env->fr_big.band = env->fr_small.band = (uint32_t)band * std::max(env->fr_big.delta, 1u) * (uint32_t)env->fr_big.dmax;
        if (const char *e = getenv("XR_FR_RAY")) env->fr_big.ray = env->fr_small.ray = std::min(FR_RAY, std::max(1, atoi(e)));
        if (const char *e = getenv("XR_FR_THREADS")) env->fr_threads_big = env->fr_threads_small = std::min(FR_T, std::max(64, atoi(e) / 32 * 32));
        if (const char *e = getenv("XR_FR_CAP")) { env->fr_big.cap_s = env->fr_small.cap_s = std::max(64, atoi(e)); env->fr_big.cap_e = env->fr_small.cap_e = std::max(64, atoi(e) / 2); }
        const int cap_g = std::max(cp, 32768), cap_ge = std::max(cp / 2, 16384);
        env->fr_big.cap_g = env->fr_small.cap_g = cap_g; env->fr_big.cap_ge = env->fr_small.cap_ge = cap_ge;
        DA(d.dist64, N * g.cells_p);
        ce = cudaMemset(d.dist64, 0xFF, sizeof(unsigned long long) * N * g.cells_p);
        if (ce != cudaSuccess) { std::string m = std::string("cudaMemset dist64: ") + cudaGetErrorString(ce); xr_free(env); return fail(nullptr, XR_E_CUDA, m); }
        DA(d.fr_epoch, N); DA(d.fr_spill, N * xr_frontier_spill_words(env->fr_big)); DA(d.minc, N * 4);
    }
    DA(d.net_win, N * (g.max_nets + 1) * 6); DA(d.fin, N); DA(d.dbg, 16 + N * 8); DA(d.netfeat, N * (g.max_nets + 1) * XR_NF);
    {   // the observation block is the big one: do not memset it twice, but report OOM clearly
        void *q = nullptr;
        ce = cudaMalloc(&q, sizeof(float) * N * (size_t)g.obs_stride);
        if (ce != cudaSuccess) {
            char buf[256];
            snprintf(buf, sizeof buf, "cudaMalloc observation buffer (%.2f GB): %s -- lower n_envs or obs_max_nets",
                     (double)(sizeof(float) * N * (size_t)g.obs_stride) / 1e9, cudaGetErrorString(ce));
            xr_free(env);
            return fail(nullptr, XR_E_CUDA, buf);
        }
        env->allocs.push_back(q);
        d.obs = reinterpret_cast<float *>(q);
    }
#undef DA
    if (cudaMallocHost(&env->p_act, sizeof(int32_t) * (5 + XR_NB * XR_NG) * N) != cudaSuccess ||
        cudaMallocHost(&env->p_flags, sizeof(int32_t) * 4) != cudaSuccess ||
        cudaMallocHost(&env->p_res, (sizeof(int32_t) * 3 + sizeof(int64_t) * XR_M_COUNT + 1) * N + 64) != cudaSuccess ||
        cudaMallocHost(&env->p_ids, sizeof(int32_t) * N) != cudaSuccess ||
        cudaMallocHost(&env->p_full, N) != cudaSuccess) {
        xr_free(env);
        return fail(nullptr, XR_E_CUDA, "cudaMallocHost failed");
    }
    env->p_lists = env->p_act + 2 * N;                  // one pinned block, same order as the device block: act | mode | grp | lists
    env->h_routed.assign(N * (g.max_nets + 1), 0);
    env->h_has_ap.assign(N * (g.max_nets + 1), 0);
    env->h_npins.assign(N * (g.max_nets + 1), 0);
    env->h_naps.assign(N * (g.max_nets + 1), 0);
    env->h_done.assign(N, 0); env->h_loaded.assign(N, 0); env->h_reset.assign(N, 0);
    env->h_nrem.assign(N, 0);
    env->h_clean.assign(N, 0);
    env->obs_mode = cfg->obs_mode == 1 ? 1 : 0;
    env->h_netwin.assign(N * (g.max_nets + 1) * 2, 0);
    env->h_area.assign(N * (g.max_nets + 1), 0);
    env->win_margin = cfg->window_margin == 0 ? 14 : cfg->window_margin;
    // 0 = auto: per step, as many CTAs per environment as keeps about two clusters per SM's worth
    env->min_cluster = cfg->min_cluster >= 16 ? 16 : cfg->min_cluster >= 8 ? 8 : cfg->min_cluster >= 4 ? 4 : cfg->min_cluster >= 2 ? 2
                     : cfg->min_cluster == 1 ? 1 : 0;
    cudaDeviceGetAttribute(&env->n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
    cudaDeviceGetAttribute(&env->smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    cudaFuncSetAttribute(k_route_win<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k_route_win2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, env->smem_cap);
    cudaFuncSetAttribute(k_route_win2<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (int k = 0; k < XR_NG; k++) {
        // the heavier the group, the higher its stream priority: its clusters are placed first
        cudaStreamCreateWithPriority(&env->gs[k], cudaStreamNonBlocking, k == 0 ? prio_lo : prio_hi);
        cudaEventCreateWithFlags(&env->ev_join[k], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&env->ev_fork, cudaEventDisableTiming);
    cudaStreamCreateWithPriority(&env->s_glob, cudaStreamNonBlocking, prio_hi);
    cudaEventCreateWithFlags(&env->ev_glob, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&env->ev_done, cudaEventDisableTiming);
    g.guide_cost = std::max(0, cfg->guide_cost); g.halo = std::max(0, std::min(8, cfg->halo));
    env->g.guide_cap = 0;
    env->engine = cfg->engine == 1 ? 1 : 0;
    if (const char *e = getenv("XR_ENGINE")) env->engine = atoi(e) == 1 ? 1 : 0;
    if ((g.guide_cost > 0 || g.halo > 0) && env->engine == 1) {
        xr_free(env);
        return fail(nullptr, XR_E_INVALID, "guide_cost / halo are implemented by the frontier engine only (engine 0)");
    }
    if (g.guide_cost > 0 || g.halo > 0) env->hybrid_area = 0;      // every net on the frontier engine
    env->metrics_mode = cfg->metrics_mode == 1 ? 1 : 0;
    if (const char *e = getenv("XR_HYBRID_AREA")) if (g.guide_cost == 0 && g.halo == 0) env->hybrid_area = std::max(0, atoi(e));
    if (const char *e = getenv("XR_HYBRID_PINS")) env->hybrid_pins = std::max(2, atoi(e));
    if (const char *e = getenv("XR_METRICS_MODE")) env->metrics_mode = atoi(e) == 1 ? 1 : 0;
    env->win_fit_cap = env->smem_cap;
    if (const char *e = getenv("XR_WIN_FIT_CAP")) env->win_fit_cap = std::min(env->smem_cap, std::max(0, atoi(e)));
    ce = xr_frontier_init(env->smem_cap);
    if (ce != cudaSuccess) { std::string m = std::string("frontier kernel attribute: ") + cudaGetErrorString(ce); xr_free(env); return fail(nullptr, XR_E_CUDA, m); }
    // tuning knobs (not part of the ABI): XR_HEAVY_PINS, XR_HEAVY_CLUSTER
    if (const char *e = getenv("XR_MEDIUM_PINS")) env->grp_pins[1] = std::max(2, atoi(e));
    if (const char *e = getenv("XR_HEAVY_PINS")) env->grp_pins[2] = std::max(env->grp_pins[1], atoi(e));
    if (const char *e = getenv("XR_HEAVY_CLUSTER")) env->heavy_cluster = atoi(e);
    if (const char *e = getenv("XR_DUAL_PINS")) env->dual_pins = atoi(e);
    if (const char *e = getenv("XR_DUAL_MINC")) env->dual_minc = std::max(2, atoi(e));
    // dynamic shared memory of the x+z sweep
    const int smem = g.Z * g.Xp * 5;
    cudaFuncSetAttribute(k_sweep_xz<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_sweep_xz<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_sweep_xz<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_sweep_xz<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_sweep_xz<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_sweep_xz<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) { std::string m = cudaGetErrorString(ce); xr_free(env); return fail(nullptr, XR_E_CUDA, m); }
    *out = env;
    return XR_OK;
}

// ------------------------------------------------------------------ load instance
extern "C" int xr_load_instance(XrEnv *env, int32_t env_id, int32_t n_block, const int32_t *block_xyz,
                                int32_t n_ap, const int32_t *ap_net, const int32_t *ap_pin,
                                const int32_t *ap_xyz) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env) return XR_E_INVALID;
    const Geo &g = env->g;
    if (env_id < 0 || env_id >= g.N) return fail(env, XR_E_INVALID, "env_id out of range");
    if (n_ap > g.max_aps) return fail(env, XR_E_CAPACITY, "n_ap exceeds max_aps");
    if (n_block < 0 || n_ap < 0 || (n_block && !block_xyz) || (n_ap && (!ap_net || !ap_pin || !ap_xyz)))
        return fail(env, XR_E_INVALID, "bad instance arrays");
    cudaSetDevice(env->device);
    env->res_on_host = false;
    // the slot holds no valid instance until this call has succeeded (the host mirrors below are rebuilt in place)
    env->h_loaded[env_id] = 0;
    env->h_reset[env_id] = 0;
    std::vector<uint32_t> ci((size_t)g.cells_p, 0);
    std::vector<uint16_t> an((size_t)g.cells_p, 0);
    for (int z = 0; z < g.Z; z++)
        for (int y = 0; y < g.Y; y++)
            for (int x = g.X; x < g.Xp; x++) ci[((size_t)z * g.Y + y) * g.Xp + x] = CI_PAD;
    auto inb = [&](int x, int y, int z) { return x >= 0 && x < g.X && y >= 0 && y < g.Y && z >= 0 && z < g.Z; };
    for (int i = 0; i < n_block; i++) {
        const int x = block_xyz[3 * i], y = block_xyz[3 * i + 1], z = block_xyz[3 * i + 2];
        if (!inb(x, y, z)) return fail(env, XR_E_INVALID, "blockage outside the grid");
        ci[((size_t)z * g.Y + y) * g.Xp + x] |= CI_BLOCK;
    }
    // sort APs by (net, pin, input order) -- same canonical order as the oracle
    std::vector<int> order(n_ap);
    for (int i = 0; i < n_ap; i++) order[i] = i;
    for (int i = 0; i < n_ap; i++) {
        if (ap_net[i] < 1 || ap_net[i] > g.max_nets) return fail(env, XR_E_CAPACITY, "net id outside 1..max_nets");
        if (ap_pin[i] < 1 || ap_pin[i] > 65535) return fail(env, XR_E_INVALID, "pin id outside 1..65535");
        if (!inb(ap_xyz[3 * i], ap_xyz[3 * i + 1], ap_xyz[3 * i + 2])) return fail(env, XR_E_INVALID, "access point outside the grid");
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (ap_net[a] != ap_net[b]) return ap_net[a] < ap_net[b];
        return ap_pin[a] < ap_pin[b];
    });
    std::vector<int32_t> cellp(g.max_aps, 0), obsoff(g.max_aps, 0), nstart(g.max_nets + 2, 0);
    std::vector<uint16_t> pin(g.max_aps, 0), srcpin(g.max_nets + 1, 0);
    std::vector<uint8_t> adj(g.max_aps, 0);
    std::vector<int32_t> netwin((size_t)(g.max_nets + 1) * 6, 0);
    for (int net = 0; net <= g.max_nets; net++) {
        env->h_netwin[((size_t)env_id * (g.max_nets + 1) + net) * 2] = 0;
        env->h_netwin[((size_t)env_id * (g.max_nets + 1) + net) * 2 + 1] = 0;
    }
    std::vector<int> sx(n_ap), sy(n_ap), sz(n_ap), snet(n_ap);
    for (int k = 0; k < n_ap; k++) {
        const int i = order[k];
        const int x = ap_xyz[3 * i], y = ap_xyz[3 * i + 1], z = ap_xyz[3 * i + 2];
        const size_t cp = ((size_t)z * g.Y + y) * g.Xp + x;
        if (an[cp] != 0) return fail(env, XR_E_INVALID, "two access points share a cell");
        if (ci[cp] & CI_BLOCK) return fail(env, XR_E_INVALID, "access point on a blockage");
        an[cp] = (uint16_t)ap_net[i];
        ci[cp] |= CI_ISAP;
        cellp[k] = (int32_t)cp; obsoff[k] = (x * g.Y + y) * g.Z + z; pin[k] = (uint16_t)ap_pin[i];
        sx[k] = x; sy[k] = y; sz[k] = z; snet[k] = ap_net[i];
        nstart[ap_net[i] + 1]++;
    }
    for (int n = 1; n <= g.max_nets + 1; n++) nstart[n] += nstart[n - 1];
    // adjacency flag (build_3Dgrid.py:126-138) and static source pin per net
    static const int DXs[6] = {1, -1, 0, 0, 0, 0}, DYs[6] = {0, 0, 1, -1, 0, 0}, DZs[6] = {0, 0, 0, 0, 1, -1};
    uint8_t *has_ap = &env->h_has_ap[(size_t)env_id * (g.max_nets + 1)];
    uint16_t *npins = &env->h_npins[(size_t)env_id * (g.max_nets + 1)];
    memset(has_ap, 0, g.max_nets + 1);
    memset(npins, 0, sizeof(uint16_t) * (g.max_nets + 1));
    std::fill(env->h_naps.begin() + (size_t)env_id * (g.max_nets + 1), env->h_naps.begin() + (size_t)(env_id + 1) * (g.max_nets + 1), 0);
    for (int k = 0; k < n_ap; k++) {
        for (int dir = 0; dir < 6; dir++) {
            const int x = sx[k] + DXs[dir], y = sy[k] + DYs[dir], z = sz[k] + DZs[dir];
            if (inb(x, y, z) && an[((size_t)z * g.Y + y) * g.Xp + x] == snet[k]) { adj[k] = 1; break; }
        }
    }
    for (int net = 1; net <= g.max_nets; net++) {
        const int s = nstart[net], t = nstart[net + 1];
        if (s == t) continue;
        has_ap[net] = 1;
        int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1, np = 0;
        for (int k = s; k < t; k++) {
            xmin = std::min(xmin, sx[k]); xmax = std::max(xmax, sx[k]);
            ymin = std::min(ymin, sy[k]); ymax = std::max(ymax, sy[k]);
            if (k == s || pin[k] != pin[k - 1]) np++;
        }
        npins[net] = (uint16_t)std::min(np, 65535);
        env->h_naps[(size_t)env_id * (g.max_nets + 1) + net] = t - s;
        env->h_area[(size_t)env_id * (g.max_nets + 1) + net] = (xmax - xmin + 1) * (ymax - ymin + 1);
        const int cx2 = xmin + xmax, cy2 = ymin + ymax;
        long best = -1; int bp = 0;
        for (int k = s; k < t; k++) {
            const long sc = labs(2L * sx[k] - cx2) + labs(2L * sy[k] - cy2);
            if (best < 0 || sc < best || (sc == best && pin[k] < bp)) { best = sc; bp = pin[k]; }
        }
        srcpin[net] = (uint16_t)bp;
        if (np - 1 > g.conn_cap) return fail(env, XR_E_CAPACITY, "net has more than 257 pins");
        // window of the on-chip search: AP bounding box + margin, clipped to the grid
        if (env->win_margin > 0) {
            const int m = env->win_margin;
            const int wx0 = std::max(0, xmin - m), wx1 = std::min(g.X - 1, xmax + m);
            const int wy0 = std::max(0, ymin - m), wy1 = std::min(g.Y - 1, ymax + m);
            int32_t *w = &netwin[(size_t)net * 6];
            w[0] = wx0 | (wy0 << 16); w[1] = (wx1 - wx0 + 1) | ((wy1 - wy0 + 1) << 16);
            w[2] = env->xc[xmin]; w[3] = env->xc[xmax]; w[4] = env->yc[ymin]; w[5] = env->yc[ymax];
            env->h_netwin[((size_t)env_id * (g.max_nets + 1) + net) * 2] = wx1 - wx0 + 1;
            env->h_netwin[((size_t)env_id * (g.max_nets + 1) + net) * 2 + 1] = wy1 - wy0 + 1;
        }
    }
    // static per-net features of the A3C observation (baseline/A3C/utils.py:236-262): half-perimeter of
    // the AP bounding box in point coordinates (x, y in DBU, z = layer), number of nets (itself
    // included) with an AP inside that box, one flag per layer holding an AP
    std::vector<float> nfeat((size_t)(g.max_nets + 1) * XR_NF, 0.f);
    {
        std::vector<int> bx0(g.max_nets + 1), bx1(g.max_nets + 1), by0(g.max_nets + 1), by1(g.max_nets + 1),
            bz0(g.max_nets + 1), bz1(g.max_nets + 1);
        for (int net = 1; net <= g.max_nets; net++) {
            const int s = nstart[net], t = nstart[net + 1];
            if (s == t) continue;
            int x0 = 1 << 30, x1 = -(1 << 30), y0 = 1 << 30, y1 = -(1 << 30), z0 = 1 << 30, z1 = -(1 << 30);
            float *nf = &nfeat[(size_t)net * XR_NF];
            for (int k = s; k < t; k++) {
                x0 = std::min(x0, env->xc[sx[k]]); x1 = std::max(x1, env->xc[sx[k]]);
                y0 = std::min(y0, env->yc[sy[k]]); y1 = std::max(y1, env->yc[sy[k]]);
                z0 = std::min(z0, sz[k]); z1 = std::max(z1, sz[k]);
                if (sz[k] < 16) nf[2 + sz[k]] = 1.f;
            }
            bx0[net] = x0; bx1[net] = x1; by0[net] = y0; by1[net] = y1; bz0[net] = z0; bz1[net] = z1;
            nf[0] = (float)((x1 - x0) + (y1 - y0) + (z1 - z0));
        }
        for (int net = 1; net <= g.max_nets; net++) {
            if (nstart[net] == nstart[net + 1]) continue;
            int conflict = 0;
            for (int o = 1; o <= g.max_nets; o++) {
                bool in = false;
                for (int k = nstart[o]; k < nstart[o + 1] && !in; k++) {
                    const int px = env->xc[sx[k]], py = env->yc[sy[k]], pz = sz[k];
                    in = px >= bx0[net] && px <= bx1[net] && py >= by0[net] && py <= by1[net] && pz >= bz0[net] && pz <= bz1[net];
                }
                conflict += in;
            }
            nfeat[(size_t)net * XR_NF + 1] = (float)conflict;
        }
    }
    const Dev &d = env->d;
    const size_t e = env_id;
    CK(cudaMemcpy(d.netfeat + e * (g.max_nets + 1) * XR_NF, nfeat.data(), sizeof(float) * nfeat.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.cellinfo + e * g.cells_p, ci.data(), sizeof(uint32_t) * g.cells_p, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.apnet + e * g.cells_p, an.data(), sizeof(uint16_t) * g.cells_p, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.ap_cellp + e * g.max_aps, cellp.data(), sizeof(int32_t) * g.max_aps, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.ap_obsoff + e * g.max_aps, obsoff.data(), sizeof(int32_t) * g.max_aps, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.ap_pin + e * g.max_aps, pin.data(), sizeof(uint16_t) * g.max_aps, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.ap_adj + e * g.max_aps, adj.data(), g.max_aps, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.net_start + e * (g.max_nets + 2), nstart.data(), sizeof(int32_t) * (g.max_nets + 2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.net_srcpin + e * (g.max_nets + 1), srcpin.data(), sizeof(uint16_t) * (g.max_nets + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.net_win + e * (g.max_nets + 1) * 6, netwin.data(), sizeof(int32_t) * (g.max_nets + 1) * 6, cudaMemcpyHostToDevice));
    env->h_loaded[env_id] = 1;
    env->h_reset[env_id] = 0;
    env->h_clean[env_id] = 0;            // the access points changed: the next reset rebuilds the observation
    return XR_OK;
}

// Route guides of environment env_id (optional cost term, XrConfig.guide_cost): boxes [n][6] = net, x0, x1, y0, y1, z in
// cells, inclusive.  Replaces the environment's previous boxes.  Synchronous.
extern "C" int xr_load_guides(XrEnv *env, int32_t env_id, int32_t n_boxes, const int32_t *boxes) {
    if (!env || env_id < 0 || env_id >= env->g.N || n_boxes < 0 || (n_boxes && !boxes)) return XR_E_INVALID;
    if (env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    Geo &g = env->g;
    cudaSetDevice(env->device);
    CK(cudaDeviceSynchronize());
    if (n_boxes > g.guide_cap || !env->d.guide_box) {               // grow the device table, keeping the other environments' boxes
        const int cap = std::max(std::max(n_boxes, 256), 2 * g.guide_cap);
        int32_t *nb = nullptr, *ns = env->d.guide_start;
        CK(cudaMalloc(&nb, sizeof(int32_t) * 5 * (size_t)cap * g.N));
        CK(cudaMemset(nb, 0, sizeof(int32_t) * 5 * (size_t)cap * g.N));
        if (!ns) {
            CK(cudaMalloc(&ns, sizeof(int32_t) * (size_t)(g.max_nets + 2) * g.N));
            CK(cudaMemset(ns, 0, sizeof(int32_t) * (size_t)(g.max_nets + 2) * g.N));
            env->allocs.push_back(ns);
            env->d.guide_start = ns;
        }
        if (env->d.guide_box) {
            for (int e = 0; e < g.N; e++)
                CK(cudaMemcpy(nb + (size_t)e * cap * 5, env->d.guide_box + (size_t)e * g.guide_cap * 5, sizeof(int32_t) * 5 * g.guide_cap, cudaMemcpyDeviceToDevice));
            env->allocs.erase(std::find(env->allocs.begin(), env->allocs.end(), (void *)env->d.guide_box));
            cudaFree(env->d.guide_box);
        }
        env->allocs.push_back(nb);
        env->d.guide_box = nb;
        g.guide_cap = cap;
    }
    std::vector<int> order(n_boxes);
    for (int i = 0; i < n_boxes; i++) {
        order[i] = i;
        if (boxes[6 * i] < 1 || boxes[6 * i] > g.max_nets) return fail(env, XR_E_INVALID, "guide box of a net outside 1..max_nets");
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return boxes[6 * a] < boxes[6 * b]; });
    std::vector<int32_t> start(g.max_nets + 2, 0), box((size_t)std::max(n_boxes, 1) * 5, 0);
    for (int k = 0; k < n_boxes; k++) {
        const int32_t *b = boxes + 6 * order[k];
        start[b[0] + 1]++;
        for (int j = 0; j < 5; j++) box[5 * (size_t)k + j] = b[1 + j];
    }
    for (int n = 1; n <= g.max_nets + 1; n++) start[n] += start[n - 1];
    CK(cudaMemcpy(env->d.guide_start + (size_t)env_id * (g.max_nets + 2), start.data(), sizeof(int32_t) * (g.max_nets + 2), cudaMemcpyHostToDevice));
    if (n_boxes) CK(cudaMemcpy(env->d.guide_box + (size_t)env_id * g.guide_cap * 5, box.data(), sizeof(int32_t) * 5 * n_boxes, cudaMemcpyHostToDevice));
    return XR_OK;
}

// ------------------------------------------------------------------------ obs
static int launch_obs(XrEnv *env, cudaStream_t st, int tag = 0) {
    const Geo &g = env->g;
    int maxn = 0;
    for (int i = 0; i < g.N; i++) maxn = std::max(maxn, std::min(env->h_nrem[i], g.obs_max_nets));
    const long long total = (2ll + 7ll * maxn) * g.cells;
    dim3 grid((unsigned)((total + OBS_CHUNK - 1) / OBS_CHUNK), g.N);
    {
        Launch L(env, XR_K_OBS, st);
        k_obs<<<grid, OBS_THREADS, 0, st>>>(env->g, env->d, tag);
    }
    CK(cudaGetLastError());
    return XR_OK;
}

static int grid_cells(const Geo &g, int per_thread) {
    const int n = (g.cells_p / per_thread + 255) / 256;
    return std::max(1, std::min(n, 4096));
}

// ---------------------------------------------------------------------- reset
extern "C" int xr_reset(XrEnv *env, const int32_t *env_ids, int32_t k, void *stream) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env) return XR_E_INVALID;
    const Geo &g = env->g;
    cudaStream_t st = (cudaStream_t)stream;
    cudaSetDevice(env->device);
    env->res_on_host = false;
    std::vector<int> ids;
    if (env_ids == nullptr) { ids.resize(g.N); for (int i = 0; i < g.N; i++) ids[i] = i; }
    else {
        if (k < 0 || k > g.N) return fail(env, XR_E_INVALID, "bad env id count");
        ids.assign(env_ids, env_ids + k);
    }
    for (int id : ids) {
        if (id < 0 || id >= g.N) return fail(env, XR_E_INVALID, "env id out of range");
        if (!env->h_loaded[id]) return fail(env, XR_E_STATE, "reset of an environment with no instance loaded");
    }
    const int nb = (g.N + 255) / 256;
    if (env_ids == nullptr) {
        Launch L(env, XR_K_MISC, st);
        k_mark<<<nb, 256, 0, st>>>(env->g, env->d, 1);
    } else {
        { Launch L(env, XR_K_MISC, st); k_mark<<<nb, 256, 0, st>>>(env->g, env->d, 0); }
        if (!ids.empty()) {
            // p_ids is reused across calls: make sure the previous upload has been consumed
            CK(cudaStreamSynchronize(st)); env->n_sync++;
            memcpy(env->p_ids, ids.data(), sizeof(int32_t) * ids.size());
            CK(cudaMemcpyAsync(env->d_ids, env->p_ids, sizeof(int32_t) * ids.size(), cudaMemcpyHostToDevice, st));
            Launch L(env, XR_K_MISC, st);
            k_mark_ids<<<((int)ids.size() + 255) / 256, 256, 0, st>>>(env->g, env->d, env->d_ids, (int)ids.size());
        }
    }
    // which environments need a full observation build (first reset, new instance, obs_mode 1)
    CK(cudaStreamSynchronize(st)); env->n_sync++;          // p_full is reused across calls
    bool any_full = false, any_inc = false;
    memset(env->p_full, 0, g.N);
    for (int id : ids) {
        const bool full = env->obs_mode == 1 || !env->h_clean[id];
        env->p_full[id] = full;
        any_full |= full; any_inc |= !full;
    }
    CK(cudaMemcpyAsync(env->d.obs_full, env->p_full, g.N, cudaMemcpyHostToDevice, st));
    if (any_inc) { Launch L(env, XR_K_OBS, st); k_obs_reset_clear<<<g.N, 256, 0, st>>>(env->g, env->d); }
    { Launch L(env, XR_K_MISC, st); k_reset_cells<<<dim3(grid_cells(g, 1), g.N), 256, 0, st>>>(env->g, env->d); }
    { Launch L(env, XR_K_MISC, st); k_reset_env<<<nb, 256, 0, st>>>(env->g, env->d); }
    if (any_inc) { Launch L(env, XR_K_OBS, st); k_obs_reset_set<<<dim3(grid_cells(g, 4), g.N), 256, 0, st>>>(env->g, env->d); }
    CK(cudaGetLastError());
    for (int id : ids) {
        uint8_t *routed = &env->h_routed[(size_t)id * (g.max_nets + 1)];
        const uint8_t *has_ap = &env->h_has_ap[(size_t)id * (g.max_nets + 1)];
        memset(routed, 0, g.max_nets + 1);
        int n = 0;
        for (int net = 1; net <= g.max_nets; net++) n += has_ap[net];
        env->h_nrem[id] = n;
        env->h_done[id] = (n == 0);
        env->h_reset[id] = 1;
        env->h_clean[id] = 1;
    }
    if (!any_full) return XR_OK;
    return launch_obs(env, st);
}

// ----------------------------------------------------------------------- step
template <int CPL>
static void launch_xz(XrEnv *env, cudaStream_t st) {
    const Geo &g = env->g;
    Launch L(env, XR_K_SWEEP_XZ, st);
    k_sweep_xz<CPL><<<dim3((g.Y + XZ_ROWS - 1) / XZ_ROWS, g.N), dim3(32, g.Z), (size_t)g.Z * g.Xp * 5, st>>>(env->g, env->d);
}
static void launch_sweep_xz(XrEnv *env, cudaStream_t st) {
    const int cpl = (env->g.Xp + 31) / 32;
    if (cpl <= 1) launch_xz<1>(env, st);
    else if (cpl <= 2) launch_xz<2>(env, st);
    else if (cpl <= 4) launch_xz<4>(env, st);
    else if (cpl <= 8) launch_xz<8>(env, st);
    else if (cpl <= 16) launch_xz<16>(env, st);
    else launch_xz<32>(env, st);
}
template <int TH>
static void launch_y(XrEnv *env, cudaStream_t st) {
    const Geo &g = env->g;
    Launch L(env, XR_K_SWEEP_Y, st);
    k_sweep_y<TH><<<dim3(g.Xp / 32, g.Z, g.N), dim3(32, (g.Y + TH - 1) / TH), 0, st>>>(env->g, env->d);
}
static void launch_sweep_y(XrEnv *env, cudaStream_t st) {
    const int Y = env->g.Y;
    if (Y <= 128) launch_y<8>(env, st);
    else if (Y <= 256) launch_y<16>(env, st);
    else if (Y <= 512) launch_y<32>(env, st);
    else launch_y<64>(env, st);
}

template <int C>
static cudaError_t launch_win_t(XrEnv *env, cudaStream_t st, int n_envs, const int *list) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_envs * C)); cfg.blockDim = dim3(WIN_T);
    cfg.dynamicSmemBytes = (size_t)env->smem_cap; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_route_win<C>, env->g, env->d, list);
}
template <int C>
static cudaError_t launch_win2_t(XrEnv *env, cudaStream_t st, int n_envs, const int *list) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_envs * C)); cfg.blockDim = dim3(WIN_T);
    cfg.dynamicSmemBytes = (size_t)env->smem_cap; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_route_win2<C>, env->g, env->d, list);
}
static int launch_route_win(XrEnv *env, cudaStream_t st, int C, bool dual, int n_envs, const int *list) {
    Launch L(env, XR_K_ROUTE_WIN, st);
    if (dual) {
        cudaError_t e = C == 2 ? launch_win2_t<2>(env, st, n_envs, list) : C == 4 ? launch_win2_t<4>(env, st, n_envs, list)
                      : C == 8 ? launch_win2_t<8>(env, st, n_envs, list) : launch_win2_t<16>(env, st, n_envs, list);
        if (e != cudaSuccess) { env->err = std::string("k_route_win2 launch: ") + cudaGetErrorString(e); return XR_E_CUDA; }
        return XR_OK;
    }
    cudaError_t e = C == 1 ? launch_win_t<1>(env, st, n_envs, list) : C == 2 ? launch_win_t<2>(env, st, n_envs, list)
                  : C == 4 ? launch_win_t<4>(env, st, n_envs, list) : C == 8 ? launch_win_t<8>(env, st, n_envs, list)
                  : launch_win_t<16>(env, st, n_envs, list);
    if (e != cudaSuccess) { env->err = std::string("k_route_win launch: ") + cudaGetErrorString(e); return XR_E_CUDA; }
    return XR_OK;
}

// A route kernel reported an error (device flag): 4 = path record overflow, 5 = open list overflow of the frontier
// engine, 1 = unreachable target, 2 / 3 = no predecessor during the walk.  The batch is left half-stepped (some
// environments finalised, the failing one partially committed): the handle refuses further steps until every
// environment has been reset (XR_E_STATE), which restores a consistent state.
static int step_failed(XrEnv *env, int code) {
    std::fill(env->h_reset.begin(), env->h_reset.end(), 0);
    env->res_on_host = false;
    if (code == 4) return fail(env, XR_E_CAPACITY, "a net's paths exceed path_capacity (XrConfig.path_capacity); reset the environments before stepping again");
    if (code == 5) return fail(env, XR_E_CAPACITY, "frontier search: open list overflow (more live entries than cells); reset the environments before stepping again");
    return fail(env, XR_E_UNROUTABLE, "maze search failed (no path / inconsistent backtrace); reset the environments before stepping again");
}

extern "C" int xr_step_async(XrEnv *env, const int32_t *actions, void *stream) {
    if (!env || !actions) return XR_E_INVALID;
    const Geo &g = env->g;
    cudaStream_t st = (cudaStream_t)stream;
    cudaSetDevice(env->device);
    if (env->pend.active) return fail(env, XR_E_STATE, "xr_step_async: the previous step has not been completed by xr_step_wait");
    // ---- validate against the host mirror of the legal sets (state untouched on error)
    bool any_route = false;
    for (int i = 0; i < g.N; i++) {
        const int a = actions[i];
        if (a == 0) continue;
        if (!env->h_reset[i]) return fail(env, XR_E_STATE, "step before reset");
        if (a == -1) continue;
        if (a < 1 || a > g.max_nets || !env->h_has_ap[(size_t)i * (g.max_nets + 1) + a] ||
            env->h_routed[(size_t)i * (g.max_nets + 1) + a] || env->h_done[i]) {
            char buf[128];
            snprintf(buf, sizeof buf, "illegal action %d for environment %d", a, i);
            return fail(env, XR_E_ILLEGAL, buf);
        }
    }
    // (p_act / p_lists are reused across calls: xr_step_wait of the previous step waited for its uploads)
    env->cur_grp = -1;
    if (env->prof) {
        cudaEvent_t e;
        if (!env->ev_pool.empty()) { e = env->ev_pool.back(); env->ev_pool.pop_back(); } else cudaEventCreate(&e);
        cudaEventRecord(e, st);
        env->cur_step_ev = e; env->step_evs.push_back(e);
    }
    static const int CS[XR_NB] = {1, 2, 4, 8, 16, 2, 4, 8, 16};
    const int NB_BAND = 5;
    int nb[XR_NG][XR_NB] = {};
    int min_cluster = env->min_cluster;
    if (min_cluster == 0) {
        int n_route = 0;
        for (int i = 0; i < g.N; i++)
            n_route += actions[i] >= 1 && env->h_npins[(size_t)i * (g.max_nets + 1) + actions[i]] >= 2;
        min_cluster = 1;
        while (min_cluster < 8 && n_route * min_cluster * 2 <= 2 * env->n_sm) min_cluster <<= 1;
    }
    bool any_global = false, any_win = false;
    int fr_grp = 0, n_routing = 0;
    for (int i = 0; i < g.N; i++) n_routing += actions[i] >= 1 && env->h_npins[(size_t)i * (g.max_nets + 1) + actions[i]] >= 2;
    const bool hybrid = env->hybrid_area > 0 && n_routing <= env->n_sm;
    std::vector<long long> fr_order;                      // frontier engine: (pin-count key, env), sorted before the upload
    int n_grp[XR_NG] = {}, n_glob[XR_NG] = {};            // environments per group / of them on the full-grid path
    int32_t *modes = env->p_lists, *grps = env->p_lists + g.N;   // then XR_NB*XR_NG lists of N: [group][bucket]
    for (int i = 0; i < g.N; i++) {
        const int a = actions[i];
        int route = 0, mode = 0, grp = 0;
        if (a >= 1 && env->h_npins[(size_t)i * (g.max_nets + 1) + a] >= 2) {
            route = a; any_route = true;
            const int np = env->h_npins[(size_t)i * (g.max_nets + 1) + a];
            const int WX = env->h_netwin[((size_t)i * (g.max_nets + 1) + a) * 2];
            const int WY = env->h_netwin[((size_t)i * (g.max_nets + 1) + a) * 2 + 1];
            int bucket = -1;
            const int mc = (np >= env->grp_pins[XR_NG - 1] && env->heavy_cluster > 0) ? env->heavy_cluster : min_cluster;
            if (env->engine == 0 && env->h_naps[(size_t)i * (g.max_nets + 1) + a] <= FR_MAXAP && np <= FR_MAXPIN) {
                // frontier engine: one CTA per net, no window.  Hybrid: when the batch is too small to fill the GPU with
                // one CTA per net, a step is as long as its largest search, and the largest searches are the few-pin
                // nets with a wide bounding box (the search floods the box on ~3 layers: the sweep kernels do that
                // faster on a cluster of CTAs).  Same results either way.
                bool wide = hybrid && np <= env->hybrid_pins && WX > 0 && WX < 1024 && WY < 1024 &&
                            env->h_area[(size_t)i * (g.max_nets + 1) + a] >= env->hybrid_area &&
                            env->h_naps[(size_t)i * (g.max_nets + 1) + a] <= WIN_TGT_CAP;
                if (wide && np > 3) {                     // nets of 4+ pins only onto a cluster: when their window fits none they
                    bool fits = false;                    // would take the full-grid sweeps in HBM, which only pays for 2-3 pins
                    for (int b = 0; b < NB_BAND && !fits; b++) {
                        if (CS[b] < mc) continue;
                        const int H = (WY + CS[b] - 1) / CS[b];
                        fits = 4ll * ((long long)g.Z * (H + 2) * (WX | 1) + WIN_AUX_WORDS(g.Z, H, WX)) <= env->win_fit_cap;
                    }
                    wide = fits;
                }
                if (!wide) {
                    fr_order.push_back(((long long)(65535 - std::min(np, 65535)) << 32) | (unsigned)i);   // many-pin nets first
                    env->n_frontier_nets++;
                    env->p_act[2 * i] = a; env->p_act[2 * i + 1] = route;
                    modes[i] = 2; grps[i] = -1;            // (group assigned below)
                    continue;
                }
            }
            if (env->engine == 0 && (g.guide_cost > 0 || g.halo > 0))
                return fail(env, XR_E_CAPACITY, "a net exceeds the frontier engine's tables (1024 access points / 264 pins), which guide_cost / halo require");
            if (WX > 0 && env->dual_pins > 0 && np >= env->dual_pins && WX < 1024 && WY < 1024 &&
                env->h_naps[(size_t)i * (g.max_nets + 1) + a] <= WIN_TGT_CAP) {
                for (int b = NB_BAND; b < XR_NB && bucket < 0; b++) {
                    if (CS[b] < mc || CS[b] < env->dual_minc) continue;
                    const long long bytes = 4ll * ((long long)WIN2_CELL_WORDS(g.Z, CS[b], WX, WY) + WIN2_AUX_WORDS(g.Z, CS[b], WX, WY));
                    if (bytes <= env->win_fit_cap) bucket = b;
                }
            }
            if (WX > 0 && bucket < 0 && WX < 1024 && WY < 1024 &&          // (the window kernels cache the net's access
                env->h_naps[(size_t)i * (g.max_nets + 1) + a] <= WIN_TGT_CAP) {   //  points on chip, packed 10+10+12 bits)
                for (int b = 0; b < NB_BAND && bucket < 0; b++) {
                    if (CS[b] < mc) continue;
                    const int H = (WY + CS[b] - 1) / CS[b];
                    const long long bytes = 4ll * ((long long)g.Z * (H + 2) * (WX | 1) + WIN_AUX_WORDS(g.Z, H, WX));
                    if (bytes <= env->win_fit_cap) bucket = b;
                }
            }
            for (int k = 1; k < XR_NG; k++) if (np >= env->grp_pins[k]) grp = k;
            if (bucket < 0) grp = XR_NG - 1;
            if (bucket >= 0) {
                mode = 1; any_win = true;
                env->p_lists[(size_t)(3 + grp * XR_NB + bucket) * g.N + nb[grp][bucket]++] = i;
                env->n_win_nets++;
            } else { any_global = true; env->n_global_nets++; n_glob[grp]++; }
        }
        if (a != 0) n_grp[grp]++;
        env->p_act[2 * i] = a; env->p_act[2 * i + 1] = route;
        modes[i] = mode; grps[i] = grp;
    }
    const int n_fr = (int)fr_order.size();
    {   // frontier nets finish in one group of their own: the heaviest group's stream when sweep kernels run beside them
        const int fg = (any_win || any_global) ? XR_NG - 1 : 0;
        for (int i = 0; i < g.N; i++) if (grps[i] < 0) { grps[i] = fg; n_grp[fg]++; }
        fr_grp = fg;
    }
    if (n_fr) {                                           // the frontier list rides in front of the bucket lists
        std::sort(fr_order.begin(), fr_order.end());
        for (int k = 0; k < n_fr; k++) env->p_lists[(size_t)2 * g.N + k] = (int32_t)(fr_order[k] & 0xFFFFFFFFll);
    }
    CK(cudaMemcpyAsync(env->d.act, env->p_act, sizeof(int32_t) * (size_t)(any_route ? (any_win || any_global ? 5 + XR_NB * XR_NG : 5) : 4) * g.N,
                       cudaMemcpyHostToDevice, st));       // act | mode | grp | env lists in one copy
    // ---- the two post-route groups run on their own streams: the light group's metric and
    // observation kernels (HBM bound) overlap the heavy group's on-chip routing
    int maxn_grp[XR_NG] = {};
    for (int i = 0; i < g.N; i++) {
        if (actions[i] == 0) continue;
        const int after = env->h_nrem[i] - (actions[i] >= 1 ? 1 : 0);
        maxn_grp[grps[i]] = std::max(maxn_grp[grps[i]], std::min(after, g.obs_max_nets));
    }
    int n_nonempty = 0;
    for (int k = 0; k < XR_NG; k++) n_nonempty += n_grp[k] > 0;
    const bool split = n_nonempty > 1 && any_route;
    // route prologue (freeze the cost flags, clear the distance field, seed the sources): once for everybody, or --
    // when the groups run on their own streams -- per group, so that the heavy group, the long pole of the step,
    // starts routing as soon as its own few environments are ready
    if (!split) {
        if (any_global) { Launch L(env, XR_K_ROUTE_BEGIN, st); k_route_begin<<<dim3(grid_cells(g, 4), g.N), 256, 0, st>>>(env->g, env->d, -1, 0); }
        { Launch L(env, XR_K_MISC, st); k_seed<<<g.N, 64, 0, st>>>(env->g, env->d, -1); }
        CK(cudaGetLastError());
        if (any_global) CK(cudaEventRecord(env->ev_glob, st));
    }
    if (split) CK(cudaEventRecord(env->ev_fork, st));
    for (int gi = 0; gi < XR_NG; gi++) {
        const int grp = XR_NG - 1 - gi;                   // heaviest group first: it is the long pole
        const bool has = n_grp[grp] > 0;
        if (!has && grp != 0) continue;                   // (group 0 always runs: it also finalises idle envs)
        cudaStream_t sg = split ? env->gs[grp] : st;
        env->cur_grp = grp;
        if (split) {
            CK(cudaStreamWaitEvent(sg, env->ev_fork, 0));
            if (n_glob[grp]) { Launch L(env, XR_K_ROUTE_BEGIN, sg); k_route_begin<<<dim3(grid_cells(g, 4), g.N), 256, 0, sg>>>(env->g, env->d, grp, 0); }
            { Launch L(env, XR_K_MISC, sg); k_seed<<<g.N, 64, 0, sg>>>(env->g, env->d, grp); }
            if (n_glob[grp]) CK(cudaEventRecord(env->ev_glob, sg));       // (only the heaviest group holds full-grid nets)
        }
        if (grp == fr_grp && n_fr) {
            const bool big = n_fr <= env->n_sm;
            Launch L(env, XR_K_ROUTE_FRONTIER, sg);
            cudaError_t e = xr_frontier_launch(env->g, env->d, env->d_lists, n_fr, big ? env->fr_big : env->fr_small,
                                               big ? env->fr_threads_big : env->fr_threads_small, sg);
            if (e != cudaSuccess) { env->err = std::string("k_route_frontier launch: ") + cudaGetErrorString(e); return XR_E_CUDA; }
        }
        for (int b = XR_NB - 1; b >= 0; b--) {              // widest clusters first: they need a whole GPC
            if (!nb[grp][b]) continue;
            int rc = launch_route_win(env, sg, CS[b], b >= NB_BAND, nb[grp][b], env->d_lists + (size_t)(1 + grp * XR_NB + b) * g.N);
            if (rc != XR_OK) return rc;
        }
        if (has && env->metrics_mode == 1) { Launch L(env, XR_K_METRICS, sg); k_metrics<<<dim3(grid_cells(g, 16), g.N), 256, 0, sg>>>(env->g, env->d, grp); }
        { Launch L(env, XR_K_MISC, sg); k_finalize<<<(g.N + 127) / 128, 128, 0, sg>>>(env->g, env->d, grp, env->metrics_mode == 0); }
        if (has) {
            Launch L(env, XR_K_OBS, sg);
            if (env->obs_mode == 1) {
                const long long total = (2ll + 7ll * maxn_grp[grp]) * g.cells;
                k_obs<<<dim3((unsigned)((total + OBS_CHUNK - 1) / OBS_CHUNK), g.N), OBS_THREADS, 0, sg>>>(env->g, env->d, grp + 2);
            } else k_obs_update<<<g.N, 256, 0, sg>>>(env->g, env->d, grp + 2);
        }
        if (split) { CK(cudaEventRecord(env->ev_join[grp], sg)); CK(cudaStreamWaitEvent(st, env->ev_join[grp], 0)); }
    }
    CK(cudaGetLastError());
    env->cur_grp = -1;
    // ---- the results ride on the same read-back as the flags, so xr_step_results needs no second round trip
    env->res_on_host = false;
    if (any_route)
        CK(cudaMemcpyAsync(env->p_res, env->d.cum, env->rb_bytes, cudaMemcpyDeviceToHost, st));   // cum | delta | flags | done
    CK(cudaEventRecord(env->ev_done, st));
    env->pend.active = true; env->pend.any_route = any_route; env->pend.any_global = any_global; env->pend.any_win = any_win;
    env->pend.st = st; env->pend.maxn = 0;
    for (int k = 0; k < XR_NG; k++) env->pend.maxn = std::max(env->pend.maxn, maxn_grp[k]);
    // ---- host mirror (the legal sets are a pure function of the actions; a step that fails on the device invalidates
    // the handle until the environments are reset, see step_failed)
    for (int i = 0; i < g.N; i++) {
        const int a = actions[i];
        if (a >= 1) {
            env->h_routed[(size_t)i * (g.max_nets + 1) + a] = 1;
            env->h_nrem[i]--;
            if (env->h_nrem[i] == 0) env->h_done[i] = 1;
        } else if (a == -1) env->h_done[i] = 1;
    }
    return XR_OK;
}

extern "C" int xr_step_wait(XrEnv *env) {
    if (!env) return XR_E_INVALID;
    if (!env->pend.active) return XR_OK;
    const Geo &g = env->g;
    cudaStream_t st = env->pend.st;
    cudaSetDevice(env->device);
    env->pend.active = false;
    // ---- full-grid sweeps, pumped from the host until every armed environment is done (flags[0] = armed environments)
    auto pump = [&](cudaStream_t sp) -> int {
        long long pumps = 0;
        const long long guard = 64ll * (g.X + g.Y + g.Z) + 4096;
        for (;;) {
            for (int p = 0; p < env->pumps_per_sync; p++) {
                launch_sweep_xz(env, sp);
                launch_sweep_y(env, sp);
                { Launch L(env, XR_K_CONTROL, sp); k_control<<<g.N, 32, 0, sp>>>(env->g, env->d); }
            }
            pumps += env->pumps_per_sync;
            if (cudaMemcpyAsync(env->p_flags, env->d.flags, sizeof(int32_t) * 2, cudaMemcpyDeviceToHost, sp) != cudaSuccess ||
                cudaStreamSynchronize(sp) != cudaSuccess) return -1;
            env->n_sync++;
            if (env->p_flags[1] != 0) return env->p_flags[1];
            if (env->p_flags[0] == 0) return 0;
            if (pumps > guard * 64) return 2;
        }
    };
    // Nets known to need the full grid (no window fits a cluster) are pumped on their own stream while the frontier and
    // window kernels of the other environments are still running: the sweeps are a chain of short dependent launches
    // that leaves most of the GPU idle.  Environments are independent; a window kernel that hands over meanwhile parks
    // its environment in phase 2, which the pumps do not touch.
    int code = 0;
    if (env->pend.any_global) {
        CK(cudaStreamWaitEvent(env->s_glob, env->ev_glob, 0));
        code = pump(env->s_glob);
    }
    CK(cudaEventSynchronize(env->ev_done)); env->n_sync++;
    bool handover = false;
    if (code == 0 && env->pend.any_route) {
        const int32_t *fl = reinterpret_cast<const int32_t *>(env->p_res + sizeof(int64_t) * XR_M_COUNT * g.N + sizeof(int32_t) * 3 * g.N);
        if (fl[1] != 0) code = fl[1];
        handover = fl[3] > 0;
    }
    if (code == 0 && handover) {
        // lazy prologue of the environments a window kernel handed over (they skipped k_route_begin): cost flags and
        // distance field over the whole grid, then the sources / the tree committed so far (which arms them).  Both
        // kernels return at once for everybody else.
        { Launch L(env, XR_K_ROUTE_BEGIN, st); k_route_begin<<<dim3(grid_cells(g, 4), g.N), 256, 0, st>>>(env->g, env->d, -1, 1); }
        { Launch L(env, XR_K_MISC, st); k_handover_seed<<<g.N, 64, 0, st>>>(env->g, env->d); }
        code = pump(st);
    }
    if (code != 0) {
        if (code < 0) return fail(env, XR_E_CUDA, "full-grid sweeps: CUDA error");
        cudaMemsetAsync(env->d.flags, 0, sizeof(int32_t) * 4, st);
        return step_failed(env, code);
    }
    if (!env->pend.any_global && !handover) env->res_on_host = env->pend.any_route;
    else {
        if (env->metrics_mode == 1) { Launch L(env, XR_K_METRICS, st); k_metrics<<<dim3(grid_cells(g, 16), g.N), 256, 0, st>>>(env->g, env->d, -1); }
        { Launch L(env, XR_K_MISC, st); k_finalize<<<(g.N + 127) / 128, 128, 0, st>>>(env->g, env->d, -1, env->metrics_mode == 0); }
        {
            const long long total = (2ll + 7ll * env->pend.maxn) * g.cells;
            Launch L(env, XR_K_OBS, st);
            if (env->obs_mode == 1)
                k_obs<<<dim3((unsigned)((total + OBS_CHUNK - 1) / OBS_CHUNK), g.N), OBS_THREADS, 0, st>>>(env->g, env->d, 1);
            else k_obs_update<<<g.N, 256, 0, st>>>(env->g, env->d, 1);
        }
        CK(cudaGetLastError());
    }
    return XR_OK;
}

extern "C" int xr_step(XrEnv *env, const int32_t *actions, void *stream) {
    const int rc = xr_step_async(env, actions, stream);
    return rc != XR_OK ? rc : xr_step_wait(env);
}

extern "C" int xr_step_results(XrEnv *env, int32_t *delta, uint8_t *done, int64_t *cum, void *stream) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env) return XR_E_INVALID;
    const Geo &g = env->g;
    cudaStream_t st = (cudaStream_t)stream;
    cudaSetDevice(env->device);
    if (env->res_on_host) {              // already read back by xr_step (nothing ran on the handle since)
        const int64_t *pc = reinterpret_cast<const int64_t *>(env->p_res);
        const int32_t *pd = reinterpret_cast<const int32_t *>(pc + (size_t)XR_M_COUNT * g.N);
        const unsigned char *pdone = reinterpret_cast<const unsigned char *>(pd + 3 * (size_t)g.N + 4);
        if (delta) memcpy(delta, pd, sizeof(int32_t) * 3 * g.N);
        if (done) memcpy(done, pdone, g.N);
        if (cum) memcpy(cum, pc, sizeof(int64_t) * XR_M_COUNT * g.N);
        return XR_OK;
    }
    if (delta) CK(cudaMemcpyAsync(delta, env->d.delta, sizeof(int32_t) * 3 * g.N, cudaMemcpyDeviceToHost, st));
    if (done) CK(cudaMemcpyAsync(done, env->d.done, g.N, cudaMemcpyDeviceToHost, st));
    if (cum) CK(cudaMemcpyAsync(cum, env->d.cum, sizeof(int64_t) * XR_M_COUNT * g.N, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); env->n_sync++;
    return XR_OK;
}

// ------------------------------------------------------------------ obs access
extern "C" int xr_obs_layout(const XrEnv *env, int64_t *obs_stride, int32_t *max_channels) {
    if (!env) return XR_E_INVALID;
    if (obs_stride) *obs_stride = env->g.obs_stride;
    if (max_channels) *max_channels = 2 + 7 * env->g.obs_max_nets;
    return XR_OK;
}
extern "C" int xr_obs_channels(const XrEnv *env, int32_t env_id, int32_t *channels) {
    if (!env || env_id < 0 || env_id >= env->g.N || !channels) return XR_E_INVALID;
    *channels = 2 + 7 * std::min(env->h_nrem[env_id], env->g.obs_max_nets);
    return XR_OK;
}
extern "C" int xr_obs_copy(XrEnv *env, int32_t env_id, float *host_out, int64_t n_floats, void *stream) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env || env_id < 0 || env_id >= env->g.N || !host_out) return XR_E_INVALID;
    const Geo &g = env->g;
    const long long need = (2ll + 7ll * std::min(env->h_nrem[env_id], g.obs_max_nets)) * g.cells;
    if (n_floats < need) return fail(env, XR_E_CAPACITY, "host observation buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    cudaSetDevice(env->device);
    CK(cudaMemcpyAsync(host_out, env->d.obs + (size_t)env_id * g.obs_stride, sizeof(float) * need, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st)); env->n_sync++;
    return XR_OK;
}

struct DlCtx { XrEnv *env; int64_t shape[5]; int64_t strides[5]; };
static void dl_deleter(XrDLManagedTensor *self) {
    DlCtx *c = reinterpret_cast<DlCtx *>(self->manager_ctx);
    xr_unref(c->env);
    delete c;
    delete self;
}
static XrDLManagedTensor *make_dl(XrEnv *env, void *data, int ndim, const int64_t *shape, const int64_t *strides,
                                  uint8_t code, uint8_t bits) {
    DlCtx *c = new DlCtx();
    c->env = env;
    for (int i = 0; i < ndim; i++) { c->shape[i] = shape[i]; c->strides[i] = strides[i]; }
    XrDLManagedTensor *m = new XrDLManagedTensor();
    m->dl_tensor.data = data;
    m->dl_tensor.device.device_type = kXrDLCUDA; m->dl_tensor.device.device_id = env->device;
    m->dl_tensor.ndim = ndim;
    m->dl_tensor.dtype.code = code; m->dl_tensor.dtype.bits = bits; m->dl_tensor.dtype.lanes = 1;
    m->dl_tensor.shape = c->shape; m->dl_tensor.strides = c->strides; m->dl_tensor.byte_offset = 0;
    m->manager_ctx = c; m->deleter = dl_deleter;
    env->refs.fetch_add(1);
    return m;
}

extern "C" int xr_obs_dlpack(XrEnv *env, int32_t env_id, void **out) {
    if (!env || !out || env_id < -1 || env_id >= env->g.N) return XR_E_INVALID;
    const Geo &g = env->g;
    const int64_t cells = g.cells;
    if (env_id >= 0) {
        const int64_t C = 2 + 7 * std::min(env->h_nrem[env_id], g.obs_max_nets);
        const int64_t shape[5] = {1, C, g.Z, g.Y, g.X};
        const int64_t strides[5] = {C * cells, cells, (int64_t)g.Y * g.X, g.X, 1};
        *out = make_dl(env, env->d.obs + (size_t)env_id * g.obs_stride, 5, shape, strides, kXrDLFloat, 32);
    } else {
        const int64_t C = 2 + 7 * g.obs_max_nets;
        const int64_t shape[5] = {g.N, C, g.Z, g.Y, g.X};
        const int64_t strides[5] = {g.obs_stride, cells, (int64_t)g.Y * g.X, g.X, 1};
        *out = make_dl(env, env->d.obs, 5, shape, strides, kXrDLFloat, 32);
    }
    return XR_OK;
}

static int buffer_info(XrEnv *env, int which, void **p, int *ndim, int64_t shape[3], uint8_t *code, uint8_t *bits) {
    const Geo &g = env->g; const Dev &d = env->d;
    shape[0] = g.N; shape[1] = 1; shape[2] = 1; *ndim = 1;
    switch (which) {
    case XR_BUF_OBS: *p = d.obs; *ndim = 2; shape[1] = g.obs_stride; *code = kXrDLFloat; *bits = 32; break;
    case XR_BUF_DELTA: *p = d.delta; *ndim = 2; shape[1] = 3; *code = kXrDLInt; *bits = 32; break;
    case XR_BUF_CUM: *p = d.cum; *ndim = 2; shape[1] = XR_M_COUNT; *code = kXrDLInt; *bits = 64; break;
    case XR_BUF_DONE: *p = d.done; *code = kXrDLUInt; *bits = 8; break;
    case XR_BUF_NREMAIN: *p = d.n_remaining; *code = kXrDLInt; *bits = 32; break;
    case XR_BUF_LEGAL: *p = d.legal; *ndim = 2; shape[1] = g.max_nets + 1; *code = kXrDLUInt; *bits = 8; break;
    case XR_BUF_STATS: *p = d.stats; shape[0] = XR_STATS_COUNT; *code = kXrDLInt; *bits = 64; break;
    case XR_BUF_REWARD: *p = d.reward; *code = kXrDLFloat; *bits = 64; break;
    case XR_BUF_NETFEAT: *p = d.netfeat; *ndim = 3; shape[1] = g.max_nets + 1; shape[2] = XR_NF; *code = kXrDLFloat; *bits = 32; break;
    default: return XR_E_INVALID;
    }
    return XR_OK;
}
extern "C" int xr_buffer_dlpack(XrEnv *env, int32_t which, void **out) {
    if (!env || !out) return XR_E_INVALID;
    void *p; int ndim; int64_t shape[3]; uint8_t code, bits;
    if (buffer_info(env, which, &p, &ndim, shape, &code, &bits) != XR_OK) return fail(env, XR_E_INVALID, "unknown buffer");
    const int64_t strides[3] = {ndim == 3 ? shape[1] * shape[2] : ndim == 2 ? shape[1] : 1, ndim == 3 ? shape[2] : 1, 1};
    *out = make_dl(env, p, ndim, shape, strides, code, bits);
    return XR_OK;
}
extern "C" int xr_buffer_ptr(XrEnv *env, int32_t which, void **dev_ptr, int64_t *n_bytes) {
    if (!env || !dev_ptr) return XR_E_INVALID;
    void *p; int ndim; int64_t shape[3]; uint8_t code, bits;
    if (buffer_info(env, which, &p, &ndim, shape, &code, &bits) != XR_OK) return fail(env, XR_E_INVALID, "unknown buffer");
    *dev_ptr = p;
    if (n_bytes) *n_bytes = shape[0] * shape[1] * shape[2] * (bits / 8);
    return XR_OK;
}

extern "C" int xr_legal_mask(const XrEnv *env, int32_t env_id, uint8_t *mask, int32_t *n_remaining) {
    if (!env || env_id < 0 || env_id >= env->g.N) return XR_E_INVALID;
    const Geo &g = env->g;
    int n = 0;
    for (int net = 0; net <= g.max_nets; net++) {
        const bool rem = net >= 1 && env->h_has_ap[(size_t)env_id * (g.max_nets + 1) + net] &&
                         !env->h_routed[(size_t)env_id * (g.max_nets + 1) + net] && !env->h_done[env_id];
        if (mask) mask[net] = rem;
        n += rem;
    }
    if (n_remaining) *n_remaining = n;
    return XR_OK;
}

// ------------------------------------------------------------- parity exports
extern "C" int xr_get_paths(XrEnv *env, int32_t env_id, int32_t *cells, int32_t cells_cap, int32_t *n_cells,
                            int32_t *conn_off, uint32_t *conn_cost, int32_t conn_cap, int32_t *n_conn) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env || env_id < 0 || env_id >= env->g.N) return XR_E_INVALID;
    const Geo &g = env->g; const Dev &d = env->d;
    cudaSetDevice(env->device);
    CK(cudaDeviceSynchronize());
    int pn = 0, cn = 0;
    CK(cudaMemcpy(&pn, d.path_n + env_id, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cn, d.conn_n + env_id, sizeof(int), cudaMemcpyDeviceToHost));
    if (n_cells) *n_cells = pn;
    if (n_conn) *n_conn = cn;
    if (pn > g.path_cap || cn > g.conn_cap) return fail(env, XR_E_CAPACITY, "path record overflowed path_capacity");
    if (cells) {
        if (cells_cap < pn) return fail(env, XR_E_CAPACITY, "cells buffer too small");
        CK(cudaMemcpy(cells, d.path + (size_t)env_id * g.path_cap, sizeof(int32_t) * pn, cudaMemcpyDeviceToHost));
    }
    if (conn_off && conn_cost) {
        if (conn_cap < cn) return fail(env, XR_E_CAPACITY, "connection buffer too small");
        CK(cudaMemcpy(conn_off, d.conn_off + (size_t)env_id * (g.conn_cap + 1), sizeof(int32_t) * (cn + 1), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(conn_cost, d.conn_cost + (size_t)env_id * g.conn_cap, sizeof(uint32_t) * cn, cudaMemcpyDeviceToHost));
    }
    return XR_OK;
}

extern "C" int xr_get_state(XrEnv *env, int32_t env_id, uint8_t *usage, uint16_t *owner) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env || env_id < 0 || env_id >= env->g.N) return XR_E_INVALID;
    const Geo &g = env->g;
    cudaSetDevice(env->device);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> ci((size_t)g.cells_p);
    CK(cudaMemcpy(ci.data(), env->d.cellinfo + (size_t)env_id * g.cells_p, sizeof(uint32_t) * g.cells_p, cudaMemcpyDeviceToHost));
    for (int z = 0; z < g.Z; z++)
        for (int y = 0; y < g.Y; y++)
            for (int x = 0; x < g.X; x++) {
                const uint32_t c = ci[((size_t)z * g.Y + y) * g.Xp + x];
                const size_t o = ((size_t)z * g.Y + y) * g.X + x;
                if (usage) usage[o] = (uint8_t)((c & CI_USAGE_MASK) >> CI_USAGE_SHIFT);
                if (owner) owner[o] = (uint16_t)(c & CI_OWNER_MASK);
            }
    return XR_OK;
}

extern "C" int xr_get_dist(XrEnv *env, int32_t env_id, uint32_t *dist) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env || env_id < 0 || env_id >= env->g.N || !dist) return XR_E_INVALID;
    const Geo &g = env->g;
    cudaSetDevice(env->device);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> dd((size_t)g.cells_p);
    CK(cudaMemcpy(dd.data(), env->d.dist + (size_t)env_id * g.cells_p, sizeof(uint32_t) * g.cells_p, cudaMemcpyDeviceToHost));
    for (int z = 0; z < g.Z; z++)
        for (int y = 0; y < g.Y; y++)
            for (int x = 0; x < g.X; x++)
                dist[((size_t)z * g.Y + y) * g.X + x] = dd[((size_t)z * g.Y + y) * g.Xp + x];
    return XR_OK;
}

// ---------------------------------------------------------------- stats / prof
extern "C" int xr_stats_update(XrEnv *env, void *stream) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env) return XR_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    cudaSetDevice(env->device);
    { Launch L(env, XR_K_MISC, st); k_stats<<<1, 256, 0, st>>>(env->g, env->d); }
    CK(cudaGetLastError());
    return XR_OK;
}

/* Multi-GPU statistics: refresh XR_BUF_STATS and sum it over the ranks of `nccl_comm` in place.  NCCL is
 * resolved at run time from the process (the library the caller created the communicator with), so the shared library
 * itself has no link-time dependency on it.                                                                           */
extern "C" int xr_stats_allreduce(XrEnv *env, void *nccl_comm, void *stream) {
    if (!env || !nccl_comm) return XR_E_INVALID;
    typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    typedef const char *(*errstr_fn)(int);
    static allreduce_fn fn = nullptr;
    static errstr_fn es = nullptr;
    if (!fn) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy already in the process (e.g. PyTorch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(env, XR_E_INVALID, std::string("xr_stats_allreduce: cannot load NCCL: ") + dlerror());
        fn = reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce"));
        es = reinterpret_cast<errstr_fn>(dlsym(h, "ncclGetErrorString"));
        if (!fn) return fail(env, XR_E_INVALID, "xr_stats_allreduce: ncclAllReduce not found");
    }
    int rc = xr_stats_update(env, stream);
    if (rc != XR_OK) return rc;
    const int nccl_int64 = 4, nccl_sum = 0;                                 // ncclDataType_t / ncclRedOp_t values of nccl.h
    const int e = fn(env->d.stats, env->d.stats, XR_STATS_COUNT, nccl_int64, nccl_sum, nccl_comm, (cudaStream_t)stream);
    if (e != 0) return fail(env, XR_E_CUDA, std::string("ncclAllReduce: ") + (es ? es(e) : "error"));
    return XR_OK;
}

extern "C" int xr_counters(const XrEnv *env, int64_t *kernel_launches, int64_t *relax_passes,
                           int64_t *cells_relaxed, int64_t *host_syncs) {
    if (!env) return XR_E_INVALID;
    const Geo &g = env->g;
    if (kernel_launches) *kernel_launches = env->n_launch;
    if (host_syncs) *host_syncs = env->n_sync;
    if (relax_passes || cells_relaxed) {
        cudaSetDevice(env->device);
        cudaDeviceSynchronize();
        std::vector<long long> es((size_t)g.N * 8);
        cudaMemcpy(es.data(), env->d.envstat, sizeof(long long) * 8 * g.N, cudaMemcpyDeviceToHost);
        long long passes = 0, cells = 0;
        for (int i = 0; i < g.N; i++) { passes += es[8 * (size_t)i + 2]; cells += es[8 * (size_t)i + 7]; }
        if (relax_passes) *relax_passes = passes;
        if (cells_relaxed) *cells_relaxed = cells;
    }
    return XR_OK;
}

extern "C" int xr_route_counters(XrEnv *env, int64_t *window_nets, int64_t *global_nets, int64_t *window_fallbacks) {
    if (!env) return XR_E_INVALID;
    cudaSetDevice(env->device);
    CK(cudaDeviceSynchronize());
    int fb = 0;
    CK(cudaMemcpy(&fb, env->d.flags + 2, sizeof(int), cudaMemcpyDeviceToHost));
    if (window_nets) *window_nets = env->n_win_nets;
    if (global_nets) *global_nets = env->n_global_nets;
    if (window_fallbacks) *window_fallbacks = fb;
    return XR_OK;
}

extern "C" int xr_frontier_counters(XrEnv *env, int64_t *frontier_nets, int64_t *rounds) {
    if (!env) return XR_E_INVALID;
    if (frontier_nets) *frontier_nets = env->n_frontier_nets;
    if (rounds) {                       // (relaxation rounds of the frontier engine = relax_passes of xr_counters)
        int64_t p = 0;
        int rc = xr_counters(env, nullptr, &p, nullptr, nullptr);
        if (rc != XR_OK) return rc;
        *rounds = p;
    }
    return XR_OK;
}

/* Stand-alone timing of one HBM-bound kernel over ALL environments of the handle (for the
 * roofline: inside a step these kernels overlap the routing of other groups).  which:
 * XR_K_OBS (rebuilds every observation in place, state unchanged) or XR_K_METRICS
 * (congestion reduction; the partial sums are discarded).  Returns the mean milliseconds of
 * `reps` launches after one warm-up, measured with CUDA events on `stream`, and the
 * algorithmic bytes one launch moves.                                                  */
extern "C" int xr_kernel_bench(XrEnv *env, int32_t which, int32_t reps, double *ms_out, double *bytes_out, void *stream) {
    if (env && env->pend.active) { const int rc__ = xr_step_wait(env); if (rc__ != XR_OK) return rc__; }
    if (!env || reps < 1 || !ms_out) return XR_E_INVALID;
    const Geo &g = env->g;
    cudaStream_t st = (cudaStream_t)stream;
    cudaSetDevice(env->device);
    for (int i = 0; i < g.N; i++) if (!env->h_reset[i]) return fail(env, XR_E_STATE, "kernel bench before reset");
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double bytes = 0;
    if (which == XR_K_OBS) {
        int maxn = 0;
        for (int i = 0; i < g.N; i++) {
            const int n = std::min(env->h_nrem[i], g.obs_max_nets);
            maxn = std::max(maxn, n);
            bytes += 4.0 * (2 + 7 * n) * g.cells;
        }
        const long long total = (2ll + 7ll * maxn) * g.cells;
        dim3 grid((unsigned)((total + OBS_CHUNK - 1) / OBS_CHUNK), g.N);
        k_mark<<<(g.N + 255) / 256, 256, 0, st>>>(env->g, env->d, 1);
        for (int r = 0; r <= reps; r++) {
            if (r == 1) cudaEventRecord(a, st);
            k_obs<<<grid, OBS_THREADS, 0, st>>>(env->g, env->d, -1);
        }
        cudaEventRecord(b, st);
        env->n_launch += reps + 2;
    } else if (which == XR_K_METRICS) {
        bytes = 4.0 * g.cells * g.N;
        for (int r = 0; r <= reps; r++) {
            if (r == 1) cudaEventRecord(a, st);
            k_metrics<<<dim3(grid_cells(g, 16), g.N), 256, 0, st>>>(env->g, env->d, -2);
        }
        cudaEventRecord(b, st);
        cudaMemsetAsync(env->d.msum, 0, sizeof(unsigned) * 4 * g.N, st);
        env->n_launch += reps + 1;
    } else { cudaEventDestroy(a); cudaEventDestroy(b); return fail(env, XR_E_INVALID, "kernel bench: unknown kernel"); }
    cudaError_t e = cudaStreamSynchronize(st);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (e != cudaSuccess) { env->err = cudaGetErrorString(e); return XR_E_CUDA; }
    *ms_out = ms / reps;
    if (bytes_out) *bytes_out = bytes;
    return XR_OK;
}

/* Profiling timeline (needs xr_profile_enable): mean offsets in ms from the start of a step
 * of, per post-route group g: out[3g+0] first route launch start, [3g+1] route end,
 * [3g+2] observation end.  Cleared on read.                                          */
extern "C" int xr_debug_timeline(XrEnv *env, double *out) {
    if (!env || !out) return XR_E_INVALID;
    cudaSetDevice(env->device);
    cudaDeviceSynchronize();
    prof_collect(env);
    for (int gidx = 0; gidx < XR_NG; gidx++)
        for (int k = 0; k < 3; k++) {
            out[3 * gidx + k] = env->tl_n[gidx][k] ? env->tl_sum[gidx][k] / env->tl_n[gidx][k] : 0.0;
            env->tl_sum[gidx][k] = 0; env->tl_n[gidx][k] = 0;
        }
    return XR_OK;
}

/* Diagnostics of the window kernel (not part of the stable ABI surface of the path):
 * out[0] iterations, [1] connections, [2] relax cycles, [3] kernel cycles (rank-0 CTAs),
 * [4] nets, [5] sum of window cells (x*y).                                            */
extern "C" int xr_debug_counters(XrEnv *env, uint64_t *out) {
    if (!env || !out) return XR_E_INVALID;
    cudaSetDevice(env->device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, env->d.dbg, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost));
    return XR_OK;
}
/* Per-environment record of the last frontier launch (kernel built with -DFR_TIMING): uint64 [N][8] = cycles, rounds,
 * expanded entries, connections, expand cycles, classify cycles, access points, largest open list. */
extern "C" int xr_debug_env_records(XrEnv *env, uint64_t *out) {
    if (!env || !out) return XR_E_INVALID;
    cudaSetDevice(env->device);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, env->d.dbg + 16, sizeof(uint64_t) * 8 * env->g.N, cudaMemcpyDeviceToHost));
    return XR_OK;
}

extern "C" int xr_profile_enable(XrEnv *env, int32_t enable) {
    if (!env) return XR_E_INVALID;
    cudaSetDevice(env->device);
    cudaDeviceSynchronize();
    prof_collect(env);
    env->prof = enable != 0;
    return XR_OK;
}
extern "C" int xr_profile_get(XrEnv *env, double *ms, int64_t *launches) {
    if (!env) return XR_E_INVALID;
    cudaSetDevice(env->device);
    cudaDeviceSynchronize();
    prof_collect(env);
    for (int k = 0; k < XR_K_COUNT; k++) {
        if (ms) ms[k] = env->prof_ms[k];
        if (launches) launches[k] = env->prof_n[k];
        env->prof_ms[k] = 0; env->prof_n[k] = 0;
    }
    return XR_OK;
}

// ------------------------------------------- stand-alone build_3Dgrid replacement
// Host side classifies the node stream exactly as getObstaclesAndAccessPoints
// (baseline/build_3Dgrid.py:6-56); the tensor itself is produced by k_obs.
extern "C" int xr_build_obs_from_nodes(int32_t device, int32_t X, int32_t Y, int32_t Z, int32_t n_nodes,
                                       const int32_t *nodes, const uint8_t *keep_nets, int32_t max_net,
                                       float *host_out, int64_t host_cap, int32_t *nets_out,
                                       int32_t *n_nets_out, void *stream) {
    if (X < 1 || Y < 1 || Z < 1 || n_nodes < 0 || (n_nodes && !nodes) || !host_out || max_net < 0)
        return fail(nullptr, XR_E_INVALID, "bad argument");
    if ((long long)X * Y * Z >= (1ll << 30)) return fail(nullptr, XR_E_INVALID, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = cudaSetDevice(device);
    if (ce != cudaSuccess) return fail(nullptr, XR_E_CUDA, cudaGetErrorString(ce));
    const int cells = X * Y * Z, cells_o = (cells + 15) / 16 * 16;
    std::vector<uint8_t> ob(cells_o, 0);
    std::vector<uint16_t> apnet(cells, 0);              // observation-layout index
    std::vector<int> cnt(max_net + 2, 0);
    auto off = [&](int x, int y, int z) { return (x * Y + y) * Z + z; };
    int n_ap = 0;
    for (int i = 0; i < n_nodes; i++) {
        const int32_t *v = nodes + 6 * (size_t)i;
        const int x = v[0], y = v[1], z = v[2], used = v[3], net = v[4];
        if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) return fail(nullptr, XR_E_INVALID, "node outside the grid");
        if (net == -1) ob[off(x, y, z)] = 1;
        else if (net == 0) { if (used == 1) ob[off(x, y, z)] = 1; }
        else {
            if (net < 1 || net > 65535) return fail(nullptr, XR_E_INVALID, "net id out of range");
            if (used == 1) ob[off(x, y, z)] = 1;
            if (net <= max_net && keep_nets && keep_nets[net]) {
                if (apnet[off(x, y, z)] == 0) { cnt[net + 1]++; n_ap++; }
                apnet[off(x, y, z)] = (uint16_t)net;
            }
        }
    }
    std::vector<int32_t> nets;
    for (int net = 1; net <= max_net; net++) if (cnt[net + 1] > 0) nets.push_back(net);
    const int n = (int)nets.size();
    const long long total = (2ll + 7ll * n) * cells;
    if (n_nets_out) *n_nets_out = n;
    if (nets_out) for (int i = 0; i < n; i++) nets_out[i] = nets[i];
    if (host_cap < total) return fail(nullptr, XR_E_CAPACITY, "host observation buffer too small");
    // AP tables
    std::vector<int32_t> nstart(max_net + 2, 0);
    for (int k = 1; k <= max_net + 1; k++) nstart[k] = nstart[k - 1] + cnt[k];
    std::vector<int32_t> fill(nstart), obsoff(std::max(n_ap, 1));
    std::vector<uint8_t> adj(std::max(n_ap, 1), 0);
    for (int x = 0; x < X; x++) for (int y = 0; y < Y; y++) for (int z = 0; z < Z; z++) {
        const int o = off(x, y, z); const int net = apnet[o];
        if (!net) continue;
        const int k = fill[net]++;
        obsoff[k] = o;
        const bool a = (x + 1 < X && apnet[off(x + 1, y, z)] == net) || (x > 0 && apnet[off(x - 1, y, z)] == net) ||
                       (y + 1 < Y && apnet[off(x, y + 1, z)] == net) || (y > 0 && apnet[off(x, y - 1, z)] == net) ||
                       (z + 1 < Z && apnet[off(x, y, z + 1)] == net) || (z > 0 && apnet[off(x, y, z - 1)] == net);
        adj[k] = a;
    }
    // one-environment device scratch
    Geo g; memset(&g, 0, sizeof g);
    g.N = 1; g.X = X; g.Y = Y; g.Z = Z; g.Xp = X; g.cells = cells; g.cells_p = cells; g.cells_o = cells_o;
    g.max_nets = std::max(max_net, 1); g.max_aps = std::max(n_ap, 1); g.obs_max_nets = g.max_nets;
    g.obs_stride = (total + 63) / 64 * 64;
    Dev d; memset(&d, 0, sizeof d);
    std::vector<void *> tmp;
    auto A = [&](size_t bytes) -> void * { void *p = nullptr; if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) return nullptr; tmp.push_back(p); return p; };
    auto freeall = [&]() { for (void *p : tmp) cudaFree(p); };
    std::vector<int32_t> rank(g.max_nets, 0), ns2(g.max_nets + 2, 0);
    for (int i = 0; i < n; i++) rank[i] = nets[i];
    for (int k = 0; k <= max_net + 1 && k < g.max_nets + 2; k++) ns2[k] = nstart[k];
    for (int k = max_net + 2; k < g.max_nets + 2; k++) ns2[k] = nstart[max_net + 1];
    uint8_t one = 1;
    d.obs = (float *)A(sizeof(float) * g.obs_stride);
    d.obst_obs = (uint8_t *)A(cells_o); d.rank_net = (int32_t *)A(sizeof(int32_t) * g.max_nets);
    d.n_remaining = (int32_t *)A(4); d.obs_do = (uint8_t *)A(1);
    d.net_start = (int32_t *)A(sizeof(int32_t) * (g.max_nets + 2));
    d.ap_obsoff = (int32_t *)A(sizeof(int32_t) * g.max_aps); d.ap_adj = (uint8_t *)A(g.max_aps);
    if (!d.obs || !d.obst_obs || !d.rank_net || !d.n_remaining || !d.obs_do || !d.net_start || !d.ap_obsoff || !d.ap_adj) {
        freeall(); return fail(nullptr, XR_E_CUDA, "cudaMalloc failed");
    }
    cudaMemcpyAsync(d.obst_obs, ob.data(), cells_o, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.rank_net, rank.data(), sizeof(int32_t) * g.max_nets, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.n_remaining, &n, 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.obs_do, &one, 1, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.net_start, ns2.data(), sizeof(int32_t) * (g.max_nets + 2), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.ap_obsoff, obsoff.data(), sizeof(int32_t) * g.max_aps, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d.ap_adj, adj.data(), g.max_aps, cudaMemcpyHostToDevice, st);
    k_obs<<<dim3((unsigned)((total + OBS_CHUNK - 1) / OBS_CHUNK), 1), OBS_THREADS, 0, st>>>(g, d, -1);
    cudaMemcpyAsync(host_out, d.obs, sizeof(float) * total, cudaMemcpyDeviceToHost, st);
    ce = cudaStreamSynchronize(st);
    freeall();
    if (ce != cudaSuccess) return fail(nullptr, XR_E_CUDA, cudaGetErrorString(ce));
    return XR_OK;
}
