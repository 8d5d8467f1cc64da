// xr_common.cuh -- shared device/host definitions of the B200 XRoute hot path.
//
// Data layout in HBM (per environment e of a batch of N):
//   router layout  : [Z][Y][Xp] x fastest, Xp = X rounded up to 32 (128-byte rows of
//                    u32), padded index cp = (z*Y + y)*Xp + x
//   observation    : float32 [C_max][cells], cell (x,y,z) at x*Y*Z + y*Z + z inside a
//                    channel (the reshape-not-permute layout of
//                    baseline/build_3Dgrid.py:97-103)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XR_INF 0x3FFFFFFFu
#define XR_ZMAX 16
#define XR_NF 22        // features per net (baseline/A3C/utils.py:212-277)

// cellinfo bits
#define CI_OWNER_MASK 0x0000FFFFu
#define CI_USAGE_SHIFT 16
#define CI_USAGE_MASK 0x00FF0000u
#define CI_BLOCK 0x01000000u
#define CI_PAD   0x02000000u
#define CI_ISAP  0x04000000u
#define CI_STATIC_MASK (CI_BLOCK | CI_PAD | CI_ISAP)

// cflag bits (per-route scratch)
#define CF_RS   1u   // another net's wire on the cell (route-shape cost)
#define CF_FS   2u   // access point of another net (fixed-shape cost)
#define CF_BLK  4u   // blockage
#define CF_TREE 8u   // cell belongs to the tree of the net being routed
#define CF_OG  16u   // outside every guide box of the net being routed (frontier engine, XrConfig.guide_cost)

// Frontier cell word (64 bit, one per cell): everything a relaxation needs in ONE load.
//   [63:35] ~epoch (29 bits)  a word of an older epoch reads as distance = infinity; all ones = older than any epoch
//   [34:5]  distance (30 bits, XR_INF = all ones)
//   [4] inside a guide box of the net being routed   [3] access point of the net being routed
//       (both set by the frontier kernel for the duration of the net)
//   [2] access point of some net   [1] blockage   [0] route shape: a committed wire covers the cell, or the cell lies in
//       the spacing halo of one (XrConfig.halo)
// atomicMin on the whole word relaxes the cell (newer epoch and smaller distance both compare smaller; the flag bits of
// a cell are the same in every candidate value).
#define FRW_RS  1ull
#define FRW_BLK 2ull
#define FRW_AP  4ull
#define FRW_OWN 8ull
#define FRW_GUIDE 16ull
#define FRW_FLAGS 31ull
#define FRW_SHIFT 5
#define FRW_STALE 0xFFFFFFFFFFFFFFE0ull
#define FRW_EPOCH_MASK 0x1FFFFFFFu
__host__ __device__ inline unsigned long long frw_make(uint32_t hi29, uint32_t dist, uint32_t flags) {
    return ((unsigned long long)hi29 << (FRW_SHIFT + 30)) | ((unsigned long long)dist << FRW_SHIFT) | flags;
}

struct Geo {
    int N, X, Y, Z, Xp;
    int cells;        // X*Y*Z
    int cells_p;      // Z*Y*Xp
    int cells_o;      // cells rounded up to 16 (stride of obst_obs)
    int max_nets, max_aps, obs_max_nets, path_cap, conn_cap;
    long long obs_stride;   // floats per environment
    int uniform_x, uniform_y, dx, dy;
    int guide_cost, halo, guide_cap;   // optional cost terms (XrConfig.guide_cost / halo); boxes per environment
    const int32_t *xc, *yc; // device copies of the track coordinates
    uint32_t multX[XR_ZMAX][4];  // 1 + GRID*[layer not horizontal] + DRC*rs + FIXED*fs
    uint32_t multY[XR_ZMAX][4];
    uint32_t multV[4];
    uint32_t pen[XR_ZMAX];       // BLOCKCOST * min_width[z] * 20
    uint32_t vlen[XR_ZMAX];      // VIACOST * pitch[z+1]: via between z and z+1
};

struct Dev {
    // static instance data
    uint32_t *cellinfo;   // [N][cells_p]
    uint16_t *apnet;      // [N][cells_p]
    int32_t  *ap_cellp;   // [N][max_aps] padded router index
    int32_t  *ap_obsoff;  // [N][max_aps] observation-layout offset
    uint16_t *ap_pin;     // [N][max_aps]
    uint8_t  *ap_adj;     // [N][max_aps] AP has a 6-neighbour AP of the same net
    int32_t  *net_start;  // [N][max_nets+2]
    uint16_t *net_srcpin; // [N][max_nets+1]
    int32_t  *net_win;    // [N][max_nets+1][6] window x0|y0<<16, WX|WY<<16, AP bbox DBU x0,x1,y0,y1
    // dynamic environment state
    uint8_t  *obst_obs;   // [N][cells_o] obstacle channel source, observation layout
    uint8_t  *routed;     // [N][max_nets+1]
    uint8_t  *legal;      // [N][max_nets+1]
    int32_t  *rank_net;   // [N][max_nets] remaining net ids ascending
    int32_t  *n_remaining;// [N]
    // route state
    uint32_t *dist;       // [N][cells_p]
    uint8_t  *cflag;      // [N][cells_p]
    int32_t  *act;        // [N][2] raw action, net to route (0 = none)
    int32_t  *mode;       // [N] 0 = global full-grid sweeps, 1 = window-resident kernel
    int32_t  *grp;        // [N] post-route group (0 light, 1 heavy): groups finish on their own streams
    int32_t  *fin;        // [N] 0 = step epilogue pending, else tag of the pass that finalised the env
    int32_t  *phase;      // [N] 0 idle, 1 routing
    int32_t  *changed;    // [N]
    int32_t  *reinit;     // [N]
    int32_t  *first;      // [N]
    uint8_t  *ap_conn;    // [N][max_aps]
    // full-grid path bookkeeping (xr_kernels_maze.cuh): dirty rows (consumed by the x+via sweep) and dirty
    // (layer, 32-column slab)s (consumed by the y sweep); "deferred" twins for relaxations skipped by the cap
    uint8_t  *g_rowd, *g_rowf;    // [N][Y]
    uint8_t  *g_slabd, *g_slabf;  // [N][Z][Xp/32]
    uint32_t *g_cap;      // [N] writes above this distance are skipped (best target distance known so far)
    uint32_t *g_gmin;     // [N] smallest distance written in the current pump
    int32_t  *g_all;      // [N] 1 = every row and slab counts as dirty in this pump (start, hand-over, re-init)
    int32_t  *flags;      // [0] envs in the global route loop, [1] error flag, [2] window fallbacks
    // results
    unsigned int *msum;   // [N][4] blocked, shorted, overflow
    int32_t  *delta;      // [N][3]
    long long *cum;       // [N][6]
    long long *wlvia;     // [N][2] cumulative wirelength / via (commit time)
    uint8_t  *done;       // [N]
    double   *reward;     // [N]
    long long *envstat;   // [N][8] steps, episodes, pumps, connections, sum dvio, sum dwl, sum dvia, cells relaxed; [2] counts relaxation passes
    long long *stats;     // [16]
    unsigned long long *dbg; // [8] window-kernel diagnostics: iterations, connections, relax cycles, kernel cycles, nets, window cells
    float    *netfeat;    // [N][max_nets+1][XR_NF] per-net feature vector (A3C flavour): HPWL, bbox conflicts, 16 layer flags | count, last d_vio, d_wl, d_via
    uint8_t  *obs_do;     // [N]
    uint8_t  *obs_full;   // [N] reset: 1 = full observation build, 0 = incremental (buffer invariant holds)
    float    *obs;        // [N][obs_stride]
    // frontier engine (xr_frontier.cu): sparse goal-directed search on an epoch-tagged field
    unsigned long long *dist64;   // [N][cells_p]  frontier cell word, see FRW_* below
    uint32_t *fr_epoch;           // [N] epoch of the last connection searched
    uint32_t *fr_spill;           // [N][4*cap_g + 2*cap_ge] open-list / expansion-list entries beyond the shared-memory part
    long long *minc;              // [N][4] blocked, shorted, overflow maintained by the commits (checked against k_metrics)
    int32_t *guide_start;         // [N][max_nets+2] guide boxes of net k: guide_box[guide_start[k] .. guide_start[k+1])
    int32_t *guide_box;           // [N][guide_cap][5] x0, x1, y0, y1, z (cells, inclusive)
    // last routed paths (parity / debug)
    int32_t  *path;       // [N][path_cap] canonical indices
    int32_t  *path_n;     // [N]
    int32_t  *conn_off;   // [N][conn_cap+1]
    uint32_t *conn_cost;  // [N][conn_cap]
    int32_t  *conn_n;     // [N]
};

__host__ __device__ inline uint32_t xr_min(uint32_t a, uint32_t b) { return a < b ? a : b; }

// weight of entering a cell with flags f along x / y on layer z, or through a via
__device__ __forceinline__ uint32_t wgt_x(const Geo &g, int z, uint32_t len, uint32_t f) {
    return len * g.multX[z][f & 3u] + ((f & CF_BLK) ? g.pen[z] : 0u);
}
__device__ __forceinline__ uint32_t wgt_y(const Geo &g, int z, uint32_t len, uint32_t f) {
    return len * g.multY[z][f & 3u] + ((f & CF_BLK) ? g.pen[z] : 0u);
}
// via between layers zl and zl+1 entering layer zv
__device__ __forceinline__ uint32_t wgt_v(const Geo &g, int zl, int zv, uint32_t f) {
    return g.vlen[zl] * g.multV[f & 3u] + ((f & CF_BLK) ? g.pen[zv] : 0u);
}

// cost flags of a cell for the net being routed, frozen from the occupancy and the access-point owner
__device__ __forceinline__ uint32_t cost_flags(uint32_t cellinfo, uint32_t apnet, uint32_t net) {
    return ((cellinfo & CI_USAGE_MASK) ? CF_RS : 0u) | ((apnet != 0u && apnet != net) ? CF_FS : 0u) |
           ((cellinfo & CI_BLOCK) ? CF_BLK : 0u);
}

__device__ __forceinline__ void dir_delta(int dir, int &ddx, int &ddy, int &ddz) {
    ddx = (dir == 0) - (dir == 1); ddy = (dir == 2) - (dir == 3); ddz = (dir == 4) - (dir == 5);
}
// weight of the move p -> c = p + delta(dir), f = flags of c (the cell entered)
__device__ __forceinline__ uint32_t move_w(const Geo &g, int px, int py, int pz, int dir, uint32_t f) {
    switch (dir) {
    case 0: return wgt_x(g, pz, (uint32_t)(g.xc[px + 1] - g.xc[px]), f);
    case 1: return wgt_x(g, pz, (uint32_t)(g.xc[px] - g.xc[px - 1]), f);
    case 2: return wgt_y(g, pz, (uint32_t)(g.yc[py + 1] - g.yc[py]), f);
    case 3: return wgt_y(g, pz, (uint32_t)(g.yc[py] - g.yc[py - 1]), f);
    case 4: return wgt_v(g, pz, pz + 1, f);
    default: return wgt_v(g, pz - 1, pz - 1, f);
    }
}
