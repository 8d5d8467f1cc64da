/*
 * xroute_b200.h -- C ABI of the B200-native XRoute environment hot path.
 *
 * One handle (XrEnv) owns a batch of N independent routing environments on one
 * GPU.  The entry points below are exactly what a binding of the reference's
 * environment boundary needs; each one names the reference interface it replaces
 * (paths relative to the xrouting/xroute_env tree):
 *
 *   reference                                              this library
 *   ------------------------------------------------------ -------------------
 *   Game.__init__ + simulator launch                        xr_create
 *     baseline/baseline_utils.py:387-390,
 *     examples/launch_training.py:56-62
 *   region dump -> first Request (nodes, nets)              xr_load_instance
 *     baseline/openroad_api/proto/net_ordering.proto:11-45
 *   Game.reset  (control socket b'initial' + first obs)     xr_reset
 *     baseline/baseline_utils.py:441-481
 *   Game.step   (Response{net_index} -> route -> Request)   xr_step
 *     baseline/baseline_utils.py:392-439
 *   Request.reward_violation / wire_length / via, deltas    xr_step_results
 *     net_ordering.proto:37-39, baseline_utils.py:426-433
 *   build_3Dgrid observation tensor                         xr_obs_dlpack /
 *     baseline/build_3Dgrid.py:224-270                      xr_obs_ptr / xr_obs_copy
 *   legal_action_set / action_space                         xr_legal_mask
 *     baseline_utils.py:438,472
 *   build_3Dgrid(data, routed_nets, bool_inference)         xr_build_obs_from_nodes
 *     baseline/build_3Dgrid.py:224-270 (eval servers,
 *     baseline/PPO/test_PPO.py:62)
 *   Game.get_feature: 22 features per net (A3C flavour)     xr_buffer_dlpack(XR_BUF_NETFEAT)
 *     baseline/A3C/utils.py:212-277
 *
 * Conventions: plain C types only; every function returns 0 on success or a
 * negative XR_E_* code (message via xr_last_error); no exceptions cross the
 * boundary; `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream).  A handle is NOT thread-safe; use one host thread per handle/GPU.
 * All device buffers are owned by the library and stay valid until xr_destroy;
 * their contents are valid until the next xr_step / xr_reset on the handle.
 */
#ifndef XROUTE_B200_H
#define XROUTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XR_VERSION 2

#define XR_OK            0
#define XR_E_INVALID    -1   /* bad argument                                    */
#define XR_E_CUDA       -2   /* CUDA runtime error (see xr_last_error)          */
#define XR_E_ILLEGAL    -3   /* illegal action (net not in the legal set); state unchanged */
#define XR_E_CAPACITY   -4   /* instance exceeds max_nets / max_aps / path capacity */
#define XR_E_UNROUTABLE -5   /* the maze search found no path (cannot happen with the default cost model) */
#define XR_E_STATE      -6   /* call sequence error (e.g. step before reset)    */

#define XR_MAX_LAYERS 16

typedef struct XrEnv XrEnv;

typedef struct XrConfig {
    int32_t device;            /* CUDA device ordinal                                        */
    int32_t n_envs;            /* N: environments in this batch (this GPU's shard)           */
    int32_t X, Y, Z;           /* track grid (Z <= XR_MAX_LAYERS, X <= 1024, Y <= 2048)      */
    int32_t max_nets;          /* largest net id any instance may use                        */
    int32_t max_aps;           /* access points per environment, upper bound                 */
    int32_t obs_max_nets;      /* nets materialised in the observation; <0 = max_nets        */
    int32_t path_capacity;     /* path cells of ONE net kept per env (xr_get_paths; the router reads them back: a net
                                  whose paths exceed it fails the step with XR_E_CAPACITY); 0 = 16(X+Y+Z)+1024 */
    const int32_t *x_coords;   /* [X] DBU, strictly increasing                               */
    const int32_t *y_coords;   /* [Y]                                                        */
    const uint8_t *layer_dir;  /* [Z] 0 = horizontal (preferred axis x), 1 = vertical        */
    const int32_t *layer_pitch;      /* [Z] DBU                                              */
    const int32_t *layer_min_width;  /* [Z] DBU                                              */
    int32_t via_cost, grid_cost, drc_cost, fixed_shape_cost, block_cost; /* router constants */
    int32_t pumps_per_sync;    /* relaxation iterations launched between host polls; 0 = default */
    int32_t window_margin;     /* cells added around a net's AP bounding box for the on-chip window
                                  search; 0 = default (14), <0 = always use the full-grid sweeps  */
    int32_t min_cluster;       /* smallest CTA cluster per environment for the window kernel
                                  (1, 2, 4 or 8); 0 = auto (from the number of routing environments) */
    int32_t obs_mode;          /* 0 = observations are updated in place each step (only the cells that
                                  change: new obstacle cells, the order channel, the access points of
                                  the net blocks that shift down) -- bit-identical to a rebuild, the
                                  consumer must treat the buffer as read-only;
                                  1 = full rebuild of every stepped environment's observation        */
    int32_t engine;            /* maze-route engine: 0 = goal-directed frontier search (default; one CTA per net on an
                                  epoch-tagged field in global memory, no window), 1 = the sweep engines (window-resident
                                  cluster kernels, full-grid sweeps as their fall-back).  Same results, bit for bit.   */
    int32_t metrics_mode;      /* 0 = blocked / shorted / overflow counts are maintained by the commits (O(path cells));
                                  1 = recomputed by a full scan of the occupancy field every step (the checker, and the
                                  HBM-bound "reward kernel" the roofline is quoted on).  Same results.                  */
    int32_t guide_cost;        /* optional cost term (run-net-ordering-training.tcl:3 -follow_guide 1, GUIDECOST 1): both
                                  multipliers gain + guide_cost on a cell outside every guide box of the net (xr_load_guides);
                                  0 = off (default).  Frontier engine only.                                             */
    int32_t halo;              /* optional cost term (SHAPEBLOATWIDTH 3.0): cells within `halo` tracks (same layer) of a routed
                                  net's wires count as route shapes (DRC cost) for the nets routed later; 0 = off (default),
                                  at most 8.  Frontier engine only.  No effect on the metrics.                          */
} XrConfig;

/* cumulative metric slots of xr_step_results / XR_BUF_CUM */
enum { XR_M_VIOLATION = 0, XR_M_WIRELENGTH = 1, XR_M_VIA = 2,
       XR_M_BLOCKED = 3, XR_M_SHORTED = 4, XR_M_OVERFLOW = 5, XR_M_COUNT = 6 };

/* device buffers exposed through xr_buffer_dlpack / xr_buffer_ptr */
enum { XR_BUF_OBS = 0,        /* float32 [N][obs_stride] (see xr_obs_layout)                 */
       XR_BUF_DELTA = 1,      /* int32   [N][3]  d_violation, d_wirelength, d_via            */
       XR_BUF_CUM = 2,        /* int64   [N][XR_M_COUNT]                                     */
       XR_BUF_DONE = 3,       /* uint8   [N]                                                 */
       XR_BUF_NREMAIN = 4,    /* int32   [N]                                                 */
       XR_BUF_LEGAL = 5,      /* uint8   [N][max_nets+1]  1 = net id still to route          */
       XR_BUF_STATS = 6,      /* int64   [XR_STATS_COUNT] per-handle sums for the all-reduce */
       XR_BUF_REWARD = 7,     /* float64 [N] -(500 dvio + 4 dvia + 0.5 dwl)                   */
       XR_BUF_NETFEAT = 8     /* float32 [N][max_nets+1][22] per-net feature vectors of the A3C flavour
                                 (baseline/A3C/utils.py:212-277): [0] half-perimeter of the AP box in point
                                 coordinates, [1] nets with an AP inside that box, [2..17] layer flags,
                                 [18] times routed since reset, [19..21] its last d_violation, d_wirelength,
                                 d_via; row 0 and rows of absent nets are zero                        */ };

/* XR_BUF_STATS slots (summed over this handle's environments; all-reduce with SUM) */
enum { XR_S_STEPS = 0, XR_S_EPISODES = 1, XR_S_VIOLATION = 2, XR_S_WIRELENGTH = 3, XR_S_VIA = 4,
       XR_S_BLOCKED = 5, XR_S_SHORTED = 6, XR_S_OVERFLOW = 7, XR_S_REWARD_X2 = 8,
       XR_S_RELAX_PASSES = 9, XR_S_CELLS_RELAXED = 10, XR_S_CONNECTIONS = 11, XR_STATS_COUNT = 16 };

/* kernel classes of xr_profile_get */
enum { XR_K_OBS = 0, XR_K_METRICS = 1, XR_K_ROUTE_BEGIN = 2, XR_K_SWEEP_XZ = 3, XR_K_SWEEP_Y = 4,
       XR_K_CONTROL = 5, XR_K_ROUTE_WIN = 6, XR_K_MISC = 7, XR_K_ROUTE_FRONTIER = 8, XR_K_COUNT = 9 };

int  xr_version(void);
int  xr_create(const XrConfig *cfg, XrEnv **out);
void xr_destroy(XrEnv *env);
const char *xr_last_error(const XrEnv *env);   /* env may be NULL: last create error */

/* Region instance for environment env_id: blockages [n_block][3] = (x,y,z); access
 * points as parallel arrays, ap_net >= 1, ap_pin >= 1, ap_xyz [n_ap][3].  AP cells
 * must be distinct and not blocked.  Synchronous (uploads the static arrays).   */
int xr_load_instance(XrEnv *env, int32_t env_id, int32_t n_block, const int32_t *block_xyz,
                     int32_t n_ap, const int32_t *ap_net, const int32_t *ap_pin,
                     const int32_t *ap_xyz);

/* Route guides of environment env_id for the optional guide term (XrConfig.guide_cost > 0): boxes [n][6] = net, x0, x1,
 * y0, y1, z in cells, inclusive (the .guide file of the design, ispd/ispd18_test1/ispd18_test1.input.guide, clipped to
 * the region).  A net without boxes has no guide term.  Replaces the previous boxes; synchronous.                     */
int xr_load_guides(XrEnv *env, int32_t env_id, int32_t n_boxes, const int32_t *boxes);

/* Reset environments env_ids[0..k) (NULL = all) to their loaded instance and
 * rebuild their observations.  Asynchronous on `stream`.                        */
int xr_reset(XrEnv *env, const int32_t *env_ids, int32_t k, void *stream);

/* One environment step for the whole batch.  actions: HOST int32 [N]; >=1 = net id to route (1-based, as Game.step),
 * 0 = leave this environment untouched, -1 = stop (marks it done).  Routes, commits, updates metrics and observations.
 * Returns XR_E_ILLEGAL (nothing changed) if any action is not legal.
 *   xr_step_async  validates on the host mirror, uploads the actions and enqueues every kernel of the step plus the
 *                  read-back of its results on `stream`; it never blocks on the device.  The actions array may be reused
 *                  as soon as it returns.  One step may be in flight per handle (XR_E_STATE otherwise).
 *   xr_step_wait   blocks until that step is complete (one event wait) and reports device-side errors.  Nets whose
 *                  search must run on the full-grid sweeps (windows that fit no cluster of CTAs, nets too large for the
 *                  on-chip tables, window searches that escaped) are finished here: that loop polls a device flag
 *                  every `pumps_per_sync` iterations, on a stream of the handle's own while the other kernels of the
 *                  step are still running.
 *   xr_step        = xr_step_async + xr_step_wait.
 * Every other entry point that touches the state (reset, results, exports) completes a pending step first.
 * A step that fails on the device (XR_E_CAPACITY, XR_E_UNROUTABLE) leaves the batch half-stepped: every environment must
 * be reset before the handle steps again (XR_E_STATE until then).                                                     */
int xr_step_async(XrEnv *env, const int32_t *actions, void *stream);
int xr_step_wait(XrEnv *env);
int xr_step(XrEnv *env, const int32_t *actions, void *stream);

/* Copy the last step's results to HOST buffers (any may be NULL) and synchronise:
 * delta int32 [N][3], done uint8 [N], cum int64 [N][XR_M_COUNT].                 */
int xr_step_results(XrEnv *env, int32_t *delta, uint8_t *done, int64_t *cum, void *stream);

/* Observation layout: per environment a float32 block of obs_stride elements, of
 * which the first (2 + 7 n_e) * X*Y*Z are the [2+7n_e, Z, Y, X] observation.     */
int xr_obs_layout(const XrEnv *env, int64_t *obs_stride, int32_t *max_channels);
int xr_obs_channels(const XrEnv *env, int32_t env_id, int32_t *channels);  /* 2 + 7 n_e (host mirror) */
int xr_obs_copy(XrEnv *env, int32_t env_id, float *host_out, int64_t n_floats, void *stream);

/* Zero-copy export.  *out receives a DLManagedTensor* (DLPack v0.8 ABI) whose
 * deleter only drops a reference; wrap it in a PyCapsule named "dltensor".
 * env_id >= 0: [1, 2+7n_e, Z, Y, X]; env_id = -1: [N, max_channels, Z, Y, X] with
 * batch stride obs_stride.                                                      */
int xr_obs_dlpack(XrEnv *env, int32_t env_id, void **out);
int xr_buffer_dlpack(XrEnv *env, int32_t which, void **out);
int xr_buffer_ptr(XrEnv *env, int32_t which, void **dev_ptr, int64_t *n_bytes);

/* Host mirror of the legal action set of one environment: mask uint8 [max_nets+1]. */
int xr_legal_mask(const XrEnv *env, int32_t env_id, uint8_t *mask, int32_t *n_remaining);

/* Parity/debug exports of one environment (synchronous).
 * xr_get_paths: cells = canonical indices (z*Y+y)*X+x of the last routed net's
 * paths, target first, connection i = cells[conn_off[i] .. conn_off[i+1]).      */
int xr_get_paths(XrEnv *env, int32_t env_id, int32_t *cells, int32_t cells_cap, int32_t *n_cells,
                 int32_t *conn_off, uint32_t *conn_cost, int32_t conn_cap, int32_t *n_conn);
int xr_get_state(XrEnv *env, int32_t env_id, uint8_t *usage, uint16_t *owner);   /* [Z][Y][X] */
int xr_get_dist(XrEnv *env, int32_t env_id, uint32_t *dist);                     /* [Z][Y][X]; the full-grid path's
                                                   scratch field: meaningful only after a net routed with window_margin = -1 */

/* Refresh XR_BUF_STATS from the per-environment counters (asynchronous).        */
int xr_stats_update(XrEnv *env, void *stream);

/* Multi-GPU: refresh XR_BUF_STATS and all-reduce it (SUM, int64 [XR_STATS_COUNT]) in place over the ranks of
 * `nccl_comm` (an ncclComm_t, passed as void*), asynchronously on `stream`.  The only collective of the path: the
 * environments of a job are sharded over the GPUs and never exchange data.  NCCL is taken
 * from the process at run time (the library the communicator was created with); no link-time dependency.      */
int xr_stats_allreduce(XrEnv *env, void *nccl_comm, void *stream);

/* Counters since creation: kernel launches issued by this library and relaxation
 * passes / cells relaxed by the maze kernels.                                   */
int xr_counters(const XrEnv *env, int64_t *kernel_launches, int64_t *relax_passes,
                int64_t *cells_relaxed, int64_t *host_syncs);
/* Route-path usage since creation: nets routed by the window kernel, nets that started
 * on the full-grid path, and window searches handed over to it (exit test failed).  */
int xr_route_counters(XrEnv *env, int64_t *window_nets, int64_t *global_nets, int64_t *window_fallbacks);
/* Nets routed by the frontier engine (XrConfig.engine 0) since creation, and the relaxation rounds they took. */
int xr_frontier_counters(XrEnv *env, int64_t *frontier_nets, int64_t *rounds);
/* Window-kernel diagnostics, uint64 [16] ([8..14] per-phase cycles when built with -DWIN_PHASE_TIMING): iterations, connections, relax cycles, kernel
 * cycles (rank-0 CTAs), nets, sum of window areas (cells per layer).                  */
int xr_debug_counters(XrEnv *env, uint64_t *out);
/* Per-environment record of the last frontier launch, uint64 [N][8] (meaningful when the library was built with
 * -DFR_TIMING): cycles, rounds, expanded entries, connections, expand cycles, classify cycles, access points,
 * largest open list.                                                                     */
int xr_debug_env_records(XrEnv *env, uint64_t *out);
/* Profiling timeline, double [9]: per post-route group (3) mean ms offsets from step start of
 * route start, route end, observation end (needs xr_profile_enable).                 */
int xr_debug_timeline(XrEnv *env, double *out);

/* Per-kernel-class device timing with CUDA events on the launching stream.
 * enable: 0/1.  xr_profile_get synchronises and returns accumulated milliseconds
 * and launch counts per XR_K_* class, then clears them.                         */
int xr_profile_enable(XrEnv *env, int32_t enable);
int xr_profile_get(XrEnv *env, double *ms /*[XR_K_COUNT]*/, int64_t *launches /*[XR_K_COUNT]*/);

/* Stand-alone timing of one HBM-bound kernel (XR_K_OBS or XR_K_METRICS) over all
 * environments, for roofline accounting: mean ms of `reps` launches (CUDA events on
 * `stream`) and the algorithmic bytes of one launch.  State is left unchanged.        */
int xr_kernel_bench(XrEnv *env, int32_t which, int32_t reps, double *ms_out, double *bytes_out, void *stream);

/* Stand-alone observation build from a decoded node stream (the eval-server path):
 * nodes int32 [n_nodes][6] = (x, y, z, used, Net, Pin) as produced by
 * handle_messange (baseline/baseline_utils.py:23-40); keep_nets uint8 [max_net+1]
 * selects the nets that stay (training: not routed; inference: in data[3]).
 * Writes the [2+7n, Z, Y, X] float32 observation to host_out (capacity in floats)
 * and the ascending net ids to nets_out.                                        */
int xr_build_obs_from_nodes(int32_t device, int32_t X, int32_t Y, int32_t Z, int32_t n_nodes,
                            const int32_t *nodes, const uint8_t *keep_nets, int32_t max_net,
                            float *host_out, int64_t host_cap, int32_t *nets_out,
                            int32_t *n_nets_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XROUTE_B200_H */
