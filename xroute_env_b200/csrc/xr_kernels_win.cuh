// xr_kernels_win.cuh -- window-resident maze route: the fast path of the router.
//
// One thread-block cluster (C = 1, 2, 4 or 8 CTAs) routes the selected net of one
// environment entirely on chip.  The search is restricted to the net's *window*
// (bounding box of all its access points plus a margin, every layer); the window is
// split into C bands of rows, each band held in the shared memory of one CTA as
// packed cells (flags << 28 | dist).  Per iteration every CTA pulls its neighbours'
// boundary rows over distributed shared memory, then runs y-, x- and via-sweeps over
// its band with one thread per line (odd row stride => bank-conflict free for both
// row- and column-wise walks); iterations repeat until no CTA of the cluster changed
// a cell.  Target choice, canonical backtrace (shared with the global path) and the
// commit follow, then the next pin -- the whole net is routed in ONE launch with no
// host round trip and no HBM traffic besides loading the flags and writing the path.
//
// Exactness: a windowed search equals the full-grid search iff no cell on the window
// boundary could lie on a path of cost <= the best target distance B found inside.
// After convergence every boundary cell c that has a grid neighbour outside the
// window is tested with d(c) + h(c) <= B, h = L1 track distance (DBU, cost >= 1 per
// DBU) from c to the bounding box of the net's access points -- an admissible lower
// bound of the cost still to pay.  If any cell passes the test the environment is
// handed to the global full-grid sweeps (xr_kernels_maze.cuh) from its current
// connection on; nothing is committed for that connection by this kernel.
#pragma once
#include <cooperative_groups.h>
#include "xr_common.cuh"
#include "xr_kernels_maze.cuh"

namespace cg = cooperative_groups;

#define WIN_T 512
#define WMASK 0x0FFFFFFFu
#define WINF 0x0FFFFFFFu
#define WBLK 0x40000000u            // packed cell: [31] tree, [30] blockage, [29] fs, [28] rs, [27:0] dist
#define WIN_TGT_CAP 256
// words of shared memory besides the cells: tables, target list, dirty flags, work list
#define WIN_AUX_WORDS(Z, H, WX)                                                                   \
    (3 * (Z) + 3 * (Z) + 24 * (Z) + (WX) + 2 + (H) + 4 + 64 + 5 * WIN_TGT_CAP +                                \
     ((Z) * (H) + (Z) * (WX) + (H) * (WX)) / 4 + 3 + ((Z) * ((H) > (WX) ? (H) : (WX))) / 2 + 2 + WIN_TGT_CAP + 8)

struct WinCtx {
    uint32_t *cell;      // [Z][HH][WXp]
    uint32_t *wlut;      // [3][Z][8] full edge weight per flag class when the axis has a uniform pitch
    uint32_t *lutm;      // [3][Z]  four 8-bit multipliers (index = rs | fs<<1) per axis and layer
    uint32_t *pens;      // [Z] blockage penalty, then [Z] via length below layer z, then [Z] above
    uint32_t *lenx;      // [WX+1]  lenx[lx] = xc[wx0+lx] - xc[wx0+lx-1]
    uint32_t *leny;      // [HH+1]  leny[ly] = yc[gy] - yc[gy-1], gy = wy0 + ry0 + ly - 1
    uint8_t *rowd, *cold, *posd;   // dirty flags: row (z,ly) / column (z,x) / via stack (ly,x)
    uint16_t *list;      // compacted dirty lines
    int *cnt;
    int Z, WX, WXp, HH, H, h;      // h = real rows of this CTA (local rows 1..h), H = band height
    int wx0, wy0, ry0;
    int uni_x, uni_y;    // uniform pitch inside the window: weights come straight from wlut
};

__device__ __forceinline__ uint32_t win_w(uint32_t lutreg, uint32_t pen, uint32_t len, uint32_t v) {
    const uint32_t m = (lutreg >> ((v >> 25) & 0x18u)) & 0xFFu;
    return len * m + ((v & WBLK) ? pen : 0u);
}

// Collect the set flags of flags[0..n) into c.list (clearing them); returns the count.
__device__ int win_compact(const WinCtx &c, uint8_t *flags, int n) {
    const int lane = threadIdx.x & 31;
    __syncthreads();                     // every thread has read the previous count and is done with the previous list
    if (threadIdx.x == 0) *c.cnt = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += WIN_T) {
        const int i = i0 + threadIdx.x;
        const bool f = i < n && flags[i];
        if (f) flags[i] = 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, f);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(c.cnt, __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (f) c.list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
        }
    }
    __syncthreads();
    return *c.cnt;
}

// Thread-per-line walk of n cells (element i at p[i*stride]) in direction DIR with carry-in
// t; cells are fetched eight at a time so the shared-memory latency overlaps.  Used when
// many lines are dirty (throughput bound: fewest instructions per cell).
#define WIN_BATCH 8
template <int DIR, bool UNI>
__device__ __forceinline__ uint32_t win_walk(uint32_t *__restrict__ p, int n, int stride, uint32_t t, uint32_t lutreg,
                                             uint32_t pen, const uint32_t *__restrict__ len,
                                             const uint32_t *__restrict__ wl,
                                             uint8_t *__restrict__ fa, int sa, uint8_t *__restrict__ fb, int sb) {
    uint32_t ch = 0xFFFFFFFFu;           // smallest distance written (0xFFFFFFFF = nothing changed)
    for (int i0 = 0; i0 < n; i0 += WIN_BATCH) {
        uint32_t v[WIN_BATCH], w[WIN_BATCH];
#pragma unroll
        for (int k = 0; k < WIN_BATCH; k++) {
            const int j = i0 + k;
            const int i = DIR > 0 ? j : n - 1 - j;
            v[k] = j < n ? p[i * stride] : 0xFFFFFFFFu;
            if (!UNI) w[k] = j < n ? len[i + (DIR < 0 ? 1 : 0)] : 0u;
        }
#pragma unroll
        for (int k = 0; k < WIN_BATCH; k++) w[k] = UNI ? wl[(v[k] >> 28) & 7u] : win_w(lutreg, pen, w[k], v[k]);
#pragma unroll
        for (int k = 0; k < WIN_BATCH; k++) {
            const int j = i0 + k;
            if (j < n) {
                const int i = DIR > 0 ? j : n - 1 - j;
                const uint32_t dcur = v[k] & WMASK;
                const uint32_t nd = xr_min(t + w[k], dcur);
                if (nd < dcur) { p[i * stride] = (v[k] & ~WMASK) | nd; fa[i * sa] = 1; fb[i * sb] = 1; ch = xr_min(ch, nd); }
                t = nd;
            }
        }
    }
    return ch;
}

// Group-cooperative exact relaxation of one line (both directions), for long lines when few
// are dirty: G lanes (a power of two, G*q >= n) share the line, lane g owns the q consecutive
// cells from g*q.  Each direction is a min-plus scan: every lane folds its cells into a
// function pair (W, D): t -> min(t + W, D), the pairs are composed across the group with
// log2(G) shuffle steps, and the carries are applied.  Serial depth 2q + log2(G) instead
// of n.  q is odd so the lanes of a group hit distinct banks.  Returns the smallest
// distance written (0xFFFFFFFF = none).
#define WIN_QMAX 7
__device__ __forceinline__ uint32_t win_scan_line(bool active, uint32_t *__restrict__ p, int n, int q, int G, int g,
                                                  uint32_t lutreg, uint32_t pen, const uint32_t *__restrict__ len,
                                                  const uint32_t *__restrict__ wl, bool uni,
                                                  uint8_t *__restrict__ fa, uint8_t *__restrict__ fb) {
    uint32_t dv[WIN_QMAX], wf[WIN_QMAX], wb[WIN_QMAX], fl[WIN_QMAX];
    const int i0 = g * q;
    unsigned chg = 0, valid = 0;
#pragma unroll
    for (int k = 0; k < WIN_QMAX; k++) {
        const int i = i0 + k;
        const bool ok = active && k < q && i < n;
        valid |= ok ? (1u << k) : 0u;
        const uint32_t v = ok ? p[i] : 0x0FFFFFFFu;
        dv[k] = v & WMASK; fl[k] = v & ~WMASK;
        if (uni) { const uint32_t w = ok ? wl[(v >> 28) & 7u] : 0u; wf[k] = w; wb[k] = w; }
        else {
            wf[k] = ok ? win_w(lutreg, pen, len[i], v) : 0u;
            wb[k] = ok ? win_w(lutreg, pen, len[i + 1], v) : 0u;
        }
    }
    {   // forward
        uint32_t W = 0, D = WINF;
#pragma unroll
        for (int k = 0; k < WIN_QMAX; k++) { D = xr_min(D + wf[k], dv[k]); W = xr_min(W + wf[k], WINF); }
        for (int off = 1; off < G; off <<= 1) {
            const uint32_t Wo = __shfl_up_sync(0xFFFFFFFFu, W, off, G);
            const uint32_t Do = __shfl_up_sync(0xFFFFFFFFu, D, off, G);
            if (g >= off) { D = xr_min(Do + W, D); W = xr_min(Wo + W, WINF); }
        }
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, D, 1, G);
        if (g == 0) t = WINF;
#pragma unroll
        for (int k = 0; k < WIN_QMAX; k++) {
            t = xr_min(t + wf[k], dv[k]);
            if (t < dv[k]) { dv[k] = t; chg |= 1u << k; }
        }
    }
    {   // backward
        uint32_t W = 0, D = WINF;
#pragma unroll
        for (int k = WIN_QMAX - 1; k >= 0; k--) { D = xr_min(D + wb[k], dv[k]); W = xr_min(W + wb[k], WINF); }
        for (int off = 1; off < G; off <<= 1) {
            const uint32_t Wo = __shfl_down_sync(0xFFFFFFFFu, W, off, G);
            const uint32_t Do = __shfl_down_sync(0xFFFFFFFFu, D, off, G);
            if (g + off < G) { D = xr_min(Do + W, D); W = xr_min(Wo + W, WINF); }
        }
        uint32_t t = __shfl_down_sync(0xFFFFFFFFu, D, 1, G);
        if (g == G - 1) t = WINF;
#pragma unroll
        for (int k = WIN_QMAX - 1; k >= 0; k--) {
            t = xr_min(t + wb[k], dv[k]);
            if (t < dv[k]) { dv[k] = t; chg |= 1u << k; }
        }
    }
    chg &= valid;                        // padding cells (identity elements) are never written
    uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < WIN_QMAX; k++) {
        if ((chg >> k) & 1u) {
            const int i = i0 + k;
            p[i] = fl[k] | dv[k];
            fa[i] = 1; fb[i] = 1;
            mn = xr_min(mn, dv[k]);
        }
    }
    return mn;
}

// one thread per dirty (z, x) column of the band: forward from the upper halo, back from the lower
__device__ uint32_t win_sweep_y(const WinCtx &c, int n, long long &work) {
    uint32_t ch = 0xFFFFFFFFu;
    for (int k = threadIdx.x; k < n; k += WIN_T) {
        const int col = c.list[k];
        const int z = col / c.WX, x = col - z * c.WX;
        uint32_t *p = c.cell + (size_t)z * c.HH * c.WXp + x;
        const uint32_t lutreg = c.lutm[1 * c.Z + z], pen = c.pens[z];
        const uint32_t *wl = c.wlut + (1 * c.Z + z) * 8;
        if (c.uni_y) {
            ch = xr_min(ch, win_walk<1, true>(p + c.WXp, c.h, c.WXp, p[0] & WMASK, lutreg, pen, c.leny + 1, wl,
                                    c.rowd + z * c.H, 1, c.posd + x, c.WX));
            ch = xr_min(ch, win_walk<-1, true>(p + c.WXp, c.h, c.WXp, p[(c.h + 1) * c.WXp] & WMASK, lutreg, pen, c.leny + 1, wl,
                                     c.rowd + z * c.H, 1, c.posd + x, c.WX));
        } else {
            ch = xr_min(ch, win_walk<1, false>(p + c.WXp, c.h, c.WXp, p[0] & WMASK, lutreg, pen, c.leny + 1, wl,
                                     c.rowd + z * c.H, 1, c.posd + x, c.WX));
            ch = xr_min(ch, win_walk<-1, false>(p + c.WXp, c.h, c.WXp, p[(c.h + 1) * c.WXp] & WMASK, lutreg, pen, c.leny + 1, wl,
                                      c.rowd + z * c.H, 1, c.posd + x, c.WX));
        }
        work += c.h;
    }
    return ch;
}

// one thread per dirty (z, ly) row of the band
#ifndef WIN_COOP_X_LINES
#define WIN_COOP_X_LINES 128         // at most this many dirty rows (measured: 96-192 best on T1-7x7 and SYN-256) ...
#endif
#ifndef WIN_COOP_X_LEN
#define WIN_COOP_X_LEN 64            // ... of at least this many cells: lanes cooperate on each row
#endif
__device__ uint32_t win_sweep_x(const WinCtx &c, int n, long long &work) {
    uint32_t ch = 0xFFFFFFFFu;
    if (n <= WIN_COOP_X_LINES && c.WX >= WIN_COOP_X_LEN && c.WX <= 32 * WIN_QMAX) {
        const int G = 32, q = ((c.WX + G - 1) / G) | 1;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int k = warp; k < n; k += WIN_T / 32) {      // one warp per row (warp-uniform trip count)
            const int row = c.list[k];
            const int z = row / c.H, ly = row - z * c.H + 1;
            uint32_t *p = c.cell + ((size_t)z * c.HH + ly) * c.WXp;
            ch = xr_min(ch, win_scan_line(true, p, c.WX, q, G, lane, c.lutm[0 * c.Z + z], c.pens[z], c.lenx,
                                          c.wlut + (0 * c.Z + z) * 8, c.uni_x != 0,
                                          c.cold + z * c.WX, c.posd + (ly - 1) * c.WX));
            if (lane == 0) work += c.WX;
        }
        return ch;
    }
    for (int k = threadIdx.x; k < n; k += WIN_T) {
        const int row = c.list[k];
        const int z = row / c.H, ly = row - z * c.H + 1;
        uint32_t *p = c.cell + ((size_t)z * c.HH + ly) * c.WXp;
        const uint32_t lutreg = c.lutm[0 * c.Z + z], pen = c.pens[z];
        const uint32_t *wl = c.wlut + (0 * c.Z + z) * 8;
        if (c.uni_x) {
            ch = xr_min(ch, win_walk<1, true>(p, c.WX, 1, WINF, lutreg, pen, c.lenx, wl, c.cold + z * c.WX, 1, c.posd + (ly - 1) * c.WX, 1));
            ch = xr_min(ch, win_walk<-1, true>(p, c.WX, 1, WINF, lutreg, pen, c.lenx, wl, c.cold + z * c.WX, 1, c.posd + (ly - 1) * c.WX, 1));
        } else {
            ch = xr_min(ch, win_walk<1, false>(p, c.WX, 1, WINF, lutreg, pen, c.lenx, wl, c.cold + z * c.WX, 1, c.posd + (ly - 1) * c.WX, 1));
            ch = xr_min(ch, win_walk<-1, false>(p, c.WX, 1, WINF, lutreg, pen, c.lenx, wl, c.cold + z * c.WX, 1, c.posd + (ly - 1) * c.WX, 1));
        }
        work += c.WX;
    }
    return ch;
}

// one thread per (ly, x) position whose via stack is dirty: up then down through the layers
// (the Z cells are fetched first so their shared-memory latency overlaps)
__device__ uint32_t win_sweep_z(const WinCtx &c, long long &work) {
    uint32_t ch = 0xFFFFFFFFu;
    const int npos = c.WX * c.h;
    const size_t zs = (size_t)c.HH * c.WXp;
    const uint32_t lutreg = c.lutm[2 * c.Z];
    for (int pos = threadIdx.x; pos < npos; pos += WIN_T) {
        if (!c.posd[pos]) continue;
        c.posd[pos] = 0;
        const int lyi = pos / c.WX, x = pos - lyi * c.WX;
        uint32_t *p = c.cell + (size_t)(lyi + 1) * c.WXp + x;
        uint32_t v[XR_ZMAX];
#pragma unroll
        for (int z = 0; z < XR_ZMAX; z++) v[z] = z < c.Z ? p[z * zs] : 0u;
        unsigned chg = 0;
        uint32_t t = v[0] & WMASK;
#pragma unroll
        for (int z = 1; z < XR_ZMAX; z++) if (z < c.Z) {
            const uint32_t dcur = v[z] & WMASK;
            t = xr_min(t + win_w(lutreg, c.pens[z], c.pens[c.Z + z], v[z]), dcur);
            if (t < dcur) { v[z] = (v[z] & ~WMASK) | t; chg |= 1u << z; ch = xr_min(ch, t); }
        }
#pragma unroll
        for (int z = XR_ZMAX - 2; z >= 0; z--) if (z < c.Z - 1) {
            const uint32_t dcur = v[z] & WMASK;
            t = xr_min(t + win_w(lutreg, c.pens[z], c.pens[2 * c.Z + z], v[z]), dcur);
            if (t < dcur) { v[z] = (v[z] & ~WMASK) | t; chg |= 1u << z; ch = xr_min(ch, t); }
        }
        if (chg) {
#pragma unroll
            for (int z = 0; z < XR_ZMAX; z++) if (z < c.Z && ((chg >> z) & 1u)) {
                p[z * zs] = v[z];
                c.rowd[z * c.H + lyi] = 1; c.cold[z * c.WX + x] = 1;
            }
        }
        work += c.Z;
    }
    return ch;
}

// Pointer to the packed cell of window coordinate (lx, wy, z), wherever in the cluster
// its band lives (generic pointer into a peer CTA's shared memory when remote).
template <int C>
__device__ __forceinline__ uint32_t *win_cell_ptr(cg::cluster_group &cluster, const WinCtx &c, int H, int lx, int wy, int z) {
    const int r = wy / H, ly = wy - r * H + 1;
    uint32_t *p = c.cell + ((size_t)z * c.HH + ly) * c.WXp + lx;
    if (C == 1) return p;
    return cluster.map_shared_rank(p, r);
}

// Mark the three lines through window cell (lx, wy, z) dirty in the CTA that owns it.
template <int C>
__device__ __forceinline__ void win_mark(cg::cluster_group &cluster, const WinCtx &c, int lx, int wy, int z) {
    const int r = wy / c.H, lyi = wy - r * c.H;
    uint8_t *rd = c.rowd + z * c.H + lyi, *cd = c.cold + z * c.WX + lx, *pd = c.posd + lyi * c.WX + lx;
    if (C > 1) { rd = cluster.map_shared_rank(rd, r); cd = cluster.map_shared_rank(cd, r); pd = cluster.map_shared_rank(pd, r); }
    *rd = 1; *cd = 1; *pd = 1;
}

template <int C>
__global__ void __launch_bounds__(WIN_T, 1) k_route_win(Geo g, Dev d, const int *env_list) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int env = env_list[blockIdx.x / C];
    const int net = d.act[2 * env + 1];
    const int tid = threadIdx.x, lane = tid & 31;
    const int *wd = d.net_win + ((size_t)env * (g.max_nets + 1) + net) * 6;
    const int wx0 = wd[0] & 0xFFFF, wy0 = wd[0] >> 16, WX = wd[1] & 0xFFFF, WY = wd[1] >> 16;
    const int H = (WY + C - 1) / C;
    WinCtx c;
    c.Z = g.Z; c.WX = WX; c.WXp = WX | 1; c.HH = H + 2; c.H = H;
    c.wx0 = wx0; c.wy0 = wy0; c.ry0 = rank * H;
    c.h = WY - c.ry0; if (c.h > H) c.h = H; if (c.h < 0) c.h = 0;
    extern __shared__ __align__(16) uint32_t wsm[];
    c.cell = wsm;
    uint32_t *aux = wsm + (size_t)c.Z * c.HH * c.WXp;
    c.lutm = aux; aux += 3 * c.Z;
    c.pens = aux; aux += 3 * c.Z;
    c.wlut = aux; aux += 24 * c.Z;
    c.uni_x = g.uniform_x; c.uni_y = g.uniform_y;
    c.lenx = aux; aux += WX + 2;
    c.leny = aux; aux += c.HH + 2;
    aux += (aux - wsm) & 1;                                 // 8-byte alignment
    unsigned long long *s_best = reinterpret_cast<unsigned long long *>(aux); aux += 4;
    int *s_flag = reinterpret_cast<int *>(aux); aux += 8;   // [0..1] changed (double buffered), [2] exit, [3] more, [4] #targets, [5] list count
    int *s_tgt = reinterpret_cast<int *>(aux); aux += 2 * WIN_TGT_CAP;   // DBU coordinates of the unconnected APs
    uint32_t *s_tloc = aux; aux += WIN_TGT_CAP;             // cell index of the unconnected APs inside this band
    uint32_t *s_red = aux; aux += 8;                        // [parity][0] smallest distance written, [1] best target in band, [4] #local targets
    // the net's access points, cached on chip (the host sends a net here only if it has at most WIN_TGT_CAP of them)
    uint32_t *s_apc = aux; aux += WIN_TGT_CAP;              // window cell: lx | wy << 10 | z << 20
    uint16_t *s_appin = reinterpret_cast<uint16_t *>(aux); aux += WIN_TGT_CAP / 2;
    uint16_t *s_tidx = reinterpret_cast<uint16_t *>(aux); aux += WIN_TGT_CAP / 2;    // unconnected access points
    uint8_t *s_apconn = reinterpret_cast<uint8_t *>(aux); aux += WIN_TGT_CAP / 4;
    uint8_t *s_apon = reinterpret_cast<uint8_t *>(aux); aux += WIN_TGT_CAP / 4;
    c.cnt = &s_flag[5];
#ifdef WIN_PHASE_CRIT
    uint32_t *s_ph = aux; aux += 8;                         // this rank's phase times of the iteration (diagnostics)
#endif
    c.rowd = reinterpret_cast<uint8_t *>(aux);
    c.cold = c.rowd + c.Z * H;
    c.posd = c.cold + c.Z * WX;
    c.list = reinterpret_cast<uint16_t *>(c.rowd + (((size_t)c.Z * H + c.Z * WX + (size_t)H * WX + 3) & ~(size_t)3));
    const int n_flag_bytes = c.Z * H + c.Z * WX + H * WX;
    // ---- tables
    for (int i = tid; i < 3 * c.Z; i += WIN_T) {
        const int axis = i / c.Z, z = i - axis * c.Z;
        uint32_t r = 0;
        for (int f = 0; f < 4; f++)
            r |= (axis == 0 ? g.multX[z][f] : axis == 1 ? g.multY[z][f] : g.multV[f]) << (8 * f);
        c.lutm[i] = r;
        c.pens[i] = axis == 0 ? g.pen[z] : axis == 1 ? (z >= 1 ? g.vlen[z - 1] : 0u) : g.vlen[z];
    }
    for (int i = tid; i < 24 * c.Z; i += WIN_T) {        // uniform-pitch weights per (axis, layer, flag class)
        const int axis = i / (8 * c.Z), z = (i / 8) % c.Z, f = i & 7;
        const uint32_t len = axis == 0 ? (uint32_t)g.dx : axis == 1 ? (uint32_t)g.dy : 0u;
        const uint32_t mult = axis == 0 ? g.multX[z][f & 3] : axis == 1 ? g.multY[z][f & 3] : g.multV[f & 3];
        c.wlut[i] = len * mult + ((f & 4) ? g.pen[z] : 0u);
    }
    for (int i = tid; i <= WX; i += WIN_T) {
        const int gx = wx0 + i;
        c.lenx[i] = (gx >= 1 && gx < g.X) ? (uint32_t)(g.xc[gx] - g.xc[gx - 1]) : 0u;
    }
    for (int i = tid; i <= c.HH; i += WIN_T) {
        const int gy = wy0 + c.ry0 + i - 1;
        c.leny[i] = (gy >= 1 && gy < g.Y) ? (uint32_t)(g.yc[gy] - g.yc[gy - 1]) : 0u;
    }
    for (int i = tid; i < n_flag_bytes; i += WIN_T) c.rowd[i] = 0;
    // ---- load the band: cost flags frozen now from the occupancy and the access-point owners, dist = INF; halo rows INF
    const size_t eoff = (size_t)env * g.cells_p;
    for (int i = tid; i < c.Z * c.HH * c.WXp; i += WIN_T) {
        const int x = i % c.WXp, ly = (i / c.WXp) % c.HH, z = i / (c.WXp * c.HH);
        uint32_t v = WINF;
        if (x < WX && ly >= 1 && ly <= c.h) {
            const int gy = wy0 + c.ry0 + ly - 1;
            const size_t gi = eoff + ((size_t)z * g.Y + gy) * g.Xp + wx0 + x;
            v |= cost_flags(__ldg(d.cellinfo + gi), __ldg(d.apnet + gi), (uint32_t)net) << 28;
        }
        c.cell[i] = v;
    }
    if (tid < 8) s_flag[tid] = 0;
    __syncthreads();
    // ---- seeds: access points of the source pin inside this band
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const int s = ns[net], t = ns[net + 1], n_ap = t - s;
    const size_t aoff = (size_t)env * g.max_aps + s;
    const unsigned srcpin = d.net_srcpin[(size_t)env * (g.max_nets + 1) + net];
    for (int i = tid; i < n_ap; i += WIN_T) {
        const int cp = d.ap_cellp[aoff + i];
        const int x = cp % g.Xp - wx0, wy = (cp / g.Xp) % g.Y - wy0, z = cp / (g.Xp * g.Y);
        const unsigned pin = d.ap_pin[aoff + i];
        s_apc[i] = (uint32_t)(x | (wy << 10) | (z << 20));
        s_tgt[2 * i] = g.xc[wx0 + x]; s_tgt[2 * i + 1] = g.yc[wy0 + wy];       // DBU position (exit-test heuristic)
        s_appin[i] = (uint16_t)pin;
        s_apconn[i] = pin == srcpin;
        if (pin != srcpin) continue;
        atomicAdd(&s_flag[6], 1);                        // every CTA counts all source APs
        const int ly = wy - c.ry0 + 1;
        if (ly >= 1 && ly <= c.h) {
            c.cell[((size_t)z * c.HH + ly) * c.WXp + x] &= ~WMASK;
            c.rowd[z * H + ly - 1] = 1; c.cold[z * WX + x] = 1; c.posd[(ly - 1) * WX + x] = 1;
        }
    }
    bool first = true;
    long long work = 0, cyc_relax = 0;
    int n_iter = 0, n_conn = 0;
#ifdef WIN_PHASE_TIMING
    long long ph[7] = {0, 0, 0, 0, 0, 0, 0};
#endif
#ifdef WIN_PHASE_CRIT
    long long phc[6] = {0, 0, 0, 0, 0, 0};
#endif
#ifdef WIN_PHASE_SPLIT
    long long sp[6] = {0, 0, 0, 0, 0, 0};     // first connection: iterations, cycles; later connections: iterations, cycles; connections; post-relax cycles
#endif
    const long long tk0 = clock64();
    const bool open_x0 = wx0 > 0, open_x1 = wx0 + WX < g.X, open_y0 = wy0 > 0, open_y1 = wy0 + WY < g.Y;
    int parity = 0;
    if (C > 1) cluster.sync(); else __syncthreads();
    const int n_src_ap = s_flag[6];
    for (;;) {                                            // ---- one connection per trip
        // ---- targets of this connection: every AP of every unconnected pin.  All CTAs keep
        // the DBU coordinates (exit-test heuristic); each keeps the cells inside its band.
        if (tid == 0) { s_flag[4] = 0; s_red[4] = 0; }
        __syncthreads();
        for (int i = tid; i < n_ap; i += WIN_T) {
            if (s_apconn[i]) continue;
            const uint32_t pc = s_apc[i];
            const int x = pc & 1023, wy = (pc >> 10) & 1023, z = pc >> 20;
            s_tidx[atomicAdd(&s_flag[4], 1)] = (uint16_t)i;
            const int ly = wy - c.ry0 + 1;
            if (ly >= 1 && ly <= c.h)
                s_tloc[atomicAdd(&s_red[4], 1u)] = (uint32_t)(((size_t)z * c.HH + ly) * c.WXp + x);
        }
        __syncthreads();
        const int n_tgt = s_flag[4];
        const int n_loc = (int)s_red[4];
        const bool early = true;                          // the target list is always complete (n_ap <= WIN_TGT_CAP)
        // ---- relax until nothing below the best target distance can change any more.
        // A sweep only ever writes values >= the value it propagates from, so once every
        // distance written in an iteration is >= B (the best target distance), all cells
        // with distance < B -- everything the target choice, the exit test and the backtrace
        // read -- are final.  The remaining dirty lines stay flagged and are folded into
        // the next connection's iterations (or dropped when the net is done).
        const long long tr0 = clock64();
#ifdef WIN_PHASE_SPLIT
        const int sp_it0 = n_iter;
#endif
        for (;;) {
            n_iter++;
            if (tid == 0) { s_red[2 * parity] = 0xFFFFFFFFu; s_red[2 * parity + 1] = 0xFFFFFFFFu; }
#ifdef WIN_PHASE_TIMING
            const long long ph0 = clock64();
#endif
            if (C > 1) {
                // pull the neighbours' boundary rows into the halo rows
                for (int i = tid; i < 2 * c.Z * WX; i += WIN_T) {
                    const int side = i / (c.Z * WX), z = (i / WX) % c.Z, x = i % WX;
                    const int nr = side == 0 ? rank - 1 : rank + 1;
                    if (nr < 0 || nr >= C) continue;
                    int nh = WY - nr * H; if (nh > H) nh = H;
                    if (nh <= 0) continue;
                    const int src_ly = side == 0 ? nh : 1;
                    const int dst_ly = side == 0 ? 0 : c.h + 1;
                    if (side == 1 && c.h < H) continue;   // no rows below a short (last) band
                    const uint32_t *rp = cluster.map_shared_rank(c.cell + ((size_t)z * c.HH + src_ly) * c.WXp + x, nr);
                    uint32_t *hp = c.cell + ((size_t)z * c.HH + dst_ly) * c.WXp + x;
                    const uint32_t nv = *rp & WMASK;
                    if (nv != (*hp & WMASK)) { *hp = nv; c.cold[z * WX + x] = 1; }
                }
                cluster.sync();
            }
#ifdef WIN_PHASE_TIMING
            const long long p0 = clock64();
#endif
            const int ny = win_compact(c, c.cold, c.Z * WX);
#ifdef WIN_PHASE_TIMING
            const long long p1 = clock64();
#endif
            uint32_t ch = win_sweep_y(c, ny, work);
#ifdef WIN_PHASE_TIMING
            __syncthreads();
            const long long p2 = clock64();
#endif
            const int nx = win_compact(c, c.rowd, c.Z * H);   // (its leading barrier closes the y sweep)
#ifdef WIN_PHASE_TIMING
            const long long p3 = clock64();
#endif
            ch = xr_min(ch, win_sweep_x(c, nx, work));
            __syncthreads();
#ifdef WIN_PHASE_TIMING
            const long long p4 = clock64();
#endif
            ch = xr_min(ch, win_sweep_z(c, work));
            ch = __reduce_min_sync(0xFFFFFFFFu, ch);
            if (lane == 0 && ch != 0xFFFFFFFFu) atomicMin(&s_red[2 * parity], ch);
            __syncthreads();                              // via sweep done: target cells are current
            if (early) {
                uint32_t bl = 0xFFFFFFFFu;
                for (int k = tid; k < n_loc; k += WIN_T) bl = xr_min(bl, c.cell[s_tloc[k]] & WMASK);
                bl = __reduce_min_sync(0xFFFFFFFFu, bl);
                if (lane == 0 && bl != 0xFFFFFFFFu) atomicMin(&s_red[2 * parity + 1], bl);
            }
#ifdef WIN_PHASE_TIMING
            const long long p5 = clock64();
            ph[0] += p1 - p0; ph[1] += p2 - p1; ph[2] += p3 - p2; ph[3] += p4 - p3; ph[4] += p5 - p4;
            ph[5] += p0 - ph0;
#endif
#ifdef WIN_PHASE_CRIT
            if (tid == 0) { s_ph[0] = (uint32_t)(p0 - ph0); s_ph[1] = (uint32_t)(p2 - p0); s_ph[2] = (uint32_t)(p4 - p2); s_ph[3] = (uint32_t)(p5 - p4); }
#endif
            if (C > 1) cluster.sync(); else __syncthreads();
#ifdef WIN_PHASE_TIMING
            ph[6] += clock64() - p5;
#endif
#ifdef WIN_PHASE_CRIT
            // critical path: per phase the slowest rank of the cluster (ph[0..3]), slowest / fastest rank total (ph[4], ph[5])
            if (tid == 0 && rank == 0) {
                uint32_t mx[4] = {0, 0, 0, 0}, tmax = 0, tmin = 0xFFFFFFFFu;
                for (int r = 0; r < C; r++) {
                    const uint32_t *q = (C > 1) ? cluster.map_shared_rank(s_ph, r) : s_ph;
                    uint32_t tot = 0;
                    for (int k = 0; k < 4; k++) { const uint32_t v = q[k]; mx[k] = v > mx[k] ? v : mx[k]; if (k) tot += v; }
                    tmax = tot > tmax ? tot : tmax; tmin = tot < tmin ? tot : tmin;
                }
                for (int k = 0; k < 4; k++) phc[k] += mx[k];
                phc[4] += tmax; phc[5] += tmin;
            }
#endif
            uint32_t gmin = 0xFFFFFFFFu, gB = 0xFFFFFFFFu;
            for (int r = 0; r < C; r++) {
                const uint32_t *rr = (C > 1) ? cluster.map_shared_rank(&s_red[2 * parity], r) : &s_red[2 * parity];
                gmin = xr_min(gmin, rr[0]); gB = xr_min(gB, rr[1]);
            }
            parity ^= 1;
            if (gmin == 0xFFFFFFFFu) break;               // fixpoint
            if (early && gB < WINF && gmin >= gB) break;  // nothing relevant can change any more
        }
        // ---- best target in this band + window-exit test
        const long long tq0 = clock64();
        cyc_relax += tq0 - tr0; n_conn++;
#ifdef WIN_PHASE_SPLIT
        if (first) { sp[0] += n_iter - sp_it0; sp[1] += tq0 - tr0; } else { sp[2] += n_iter - sp_it0; sp[3] += tq0 - tr0; }
        sp[4] += 1;
#endif
        if (tid == 0) s_best[0] = ~0ull;
        __syncthreads();
        unsigned long long best = ~0ull;
        for (int k = tid; k < n_tgt; k += WIN_T) {
            const uint32_t pc = s_apc[s_tidx[k]];
            const int x = pc & 1023, wy = (pc >> 10) & 1023, z = pc >> 20;
            const int ly = wy - c.ry0 + 1;
            if (ly < 1 || ly > c.h) continue;
            const uint32_t dv = c.cell[((size_t)z * c.HH + ly) * c.WXp + x] & WMASK;
            const unsigned cp = (unsigned)((z * g.Y + wy0 + wy) * g.Xp + wx0 + x);
            const unsigned long long key = ((unsigned long long)dv << 32) | cp;
            best = key < best ? key : best;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, best, off);
            best = o < best ? o : best;
        }
        if (lane == 0 && best != ~0ull) atomicMin(&s_best[0], best);
        if (C > 1) cluster.sync(); else __syncthreads();
        best = ~0ull;
        for (int r = 0; r < C; r++) {
            const unsigned long long o = (C > 1) ? *cluster.map_shared_rank(&s_best[0], r) : s_best[0];
            best = o < best ? o : best;
        }
        const uint32_t B = (uint32_t)(best >> 32);
        // exit test over the open faces of the band
        bool esc = (best == ~0ull) || B >= WINF;
        if (!esc && (open_x0 || open_x1 || open_y0 || open_y1)) {
            const int nface = c.Z * c.h * 2 + c.Z * WX * 2;
            for (int i = tid; i < nface && !esc; i += WIN_T) {
                int x, ly, z; bool open;
                if (i < c.Z * c.h * 2) {
                    const int side = i & 1, k = i >> 1; z = k / c.h; ly = k - z * c.h + 1;
                    x = side ? WX - 1 : 0; open = side ? open_x1 : open_x0;
                } else {
                    const int k0 = i - c.Z * c.h * 2; const int side = k0 & 1, k = k0 >> 1; z = k / WX; x = k - z * WX;
                    if (side == 0) { ly = 1; open = open_y0 && rank == 0; }
                    else { ly = c.h; open = open_y1 && (c.ry0 + c.h == WY); }
                    if (c.h == 0) open = false;
                }
                if (!open) continue;
                const uint32_t dv = c.cell[((size_t)z * c.HH + ly) * c.WXp + x] & WMASK;
                if (dv >= WINF) continue;
                if (dv > B) continue;
                const int px = g.xc[wx0 + x], py = g.yc[wy0 + c.ry0 + ly - 1];
                // admissible remaining cost: L1 track distance (>= 1 cost unit per DBU) to the
                // nearest unconnected access point; bounding box of all APs if the list overflowed
                uint32_t hmin = 0xFFFFFFFFu;
                for (int j = 0; j < n_tgt; j++) {
                    const int a = s_tidx[j];
                    const uint32_t hh = (uint32_t)(abs(px - s_tgt[2 * a]) + abs(py - s_tgt[2 * a + 1]));
                    hmin = hh < hmin ? hh : hmin;
                }
                if (dv + hmin <= B) esc = true;
            }
        }
        const int esc_any = __syncthreads_or(esc);
        if (C > 1) {
            if (tid == 0) s_flag[2] = esc_any;
            cluster.sync();
            int tot = 0;
            for (int r = 0; r < C; r++) tot |= *cluster.map_shared_rank(&s_flag[2], r);
            if (tot) {
                // hand the environment to the global path (rank 0 arms it); all CTAs leave
                if (rank == 0 && tid == 0) {
                    d.phase[env] = 2; d.changed[env] = 1; d.reinit[env] = first ? 0 : 1; d.first[env] = first ? 1 : 0;
                    atomicAdd(&d.flags[3], 1); atomicAdd(&d.flags[2], 1);
                }
                break;
            }
        } else if (esc_any) {
            if (tid == 0) {
                d.phase[env] = 2; d.changed[env] = 1; d.reinit[env] = first ? 0 : 1; d.first[env] = first ? 1 : 0;
                atomicAdd(&d.flags[3], 1); atomicAdd(&d.flags[2], 1);
            }
            break;
        }
        // ---- canonical backtrace + commit by warp 0 of rank 0 (cells read over DSMEM)
        if (rank == 0 && tid < 32) {
            int cp = (int)(best & 0xFFFFFFFFu);
            int cx = cp % g.Xp, cy = (cp / g.Xp) % g.Y, cz = cp / (g.Xp * g.Y);
            int pn = d.path_n[env];
            const int pn0 = pn;
            const int cn = d.conn_n[env];
            int *path = d.path + (size_t)env * g.path_cap;
            long long wl = 0, via = 0;
            int last = -1;
            bool fail = false;
            auto inwin = [&](int x, int y, int z) {
                return x >= wx0 && x < wx0 + WX && y >= wy0 && y < wy0 + WY && z >= 0 && z < g.Z;
            };
            // The walk only touches the on-chip cells and the path list; the global state of the new path cells
            // (occupancy, observation bytes, tree flags) is committed from a pending list (s_tloc is free here) in
            // parallel passes instead of one global round trip per straight run.
            int pend = 0;
            auto flush = [&]() {
                __syncwarp();
                for (int k = lane; k < pend; k += 32) {
                    const uint32_t q = s_tloc[k];
                    commit_cell(g, d, env, net, wx0 + (int)(q & 1023u), wy0 + (int)((q >> 10) & 1023u), (int)(q >> 20));
                }
                __syncwarp();
                pend = 0;
            };
            for (;;) {
                __syncwarp();
                uint32_t *pc = win_cell_ptr<C>(cluster, c, H, cx - wx0, cy - wy0, cz);
                const uint32_t vc = *pc;
                const uint32_t dc = vc & WMASK;
                if (dc == 0) break;
                if (last >= 0) {
                    int ddx, ddy, ddz; dir_delta(last, ddx, ddy, ddz);
                    const int ax = cx - lane * ddx, ay = cy - lane * ddy, az = cz - lane * ddz;
                    const int bx = ax - ddx, by = ay - ddy, bz = az - ddz;
                    bool ok = inwin(ax, ay, az) && inwin(bx, by, bz);
                    uint32_t *pa = nullptr;
                    if (ok) {
                        pa = win_cell_ptr<C>(cluster, c, H, ax - wx0, ay - wy0, az);
                        const uint32_t va = *pa, vb = *win_cell_ptr<C>(cluster, c, H, bx - wx0, by - wy0, bz);
                        const uint32_t da = va & WMASK, db = vb & WMASK;
                        ok = da != 0 && db < WINF && db + move_w(g, bx, by, bz, last, va >> 28) == da;
                    }
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                    const int run = (m == 0xFFFFFFFFu) ? 32 : (__ffs(~m) - 1);
                    if (run > 0) {
                        if (pend + 32 > WIN_TGT_CAP) flush();
                        if (lane < run) {
                            s_tloc[pend + lane] = (uint32_t)((ax - wx0) | ((ay - wy0) << 10) | (az << 20));
                            if (pn + lane < g.path_cap) path[pn + lane] = (az * g.Y + ay) * g.X + ax;
                            if (last >= 4) via += 1;
                            else if (last < 2) wl += abs(g.xc[ax] - g.xc[bx]);
                            else wl += abs(g.yc[ay] - g.yc[by]);
                        }
                        pn += run; pend += run;
                        cx -= run * ddx; cy -= run * ddy; cz -= run * ddz;
                        continue;
                    }
                }
                bool ok = false;
                int px = 0, py = 0, pz = 0;
                if (lane < 6) {
                    int ddx, ddy, ddz; dir_delta(lane, ddx, ddy, ddz);
                    px = cx - ddx; py = cy - ddy; pz = cz - ddz;
                    if (inwin(px, py, pz)) {
                        const uint32_t dp = *win_cell_ptr<C>(cluster, c, H, px - wx0, py - wy0, pz) & WMASK;
                        ok = dp < WINF && dp + move_w(g, px, py, pz, lane, vc >> 28) == dc;
                    }
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                if (m == 0u) { fail = true; break; }
                const int dir = __ffs(m) - 1;
                if (pend + 1 > WIN_TGT_CAP) flush();
                if (lane == dir) {
                    s_tloc[pend] = (uint32_t)((cx - wx0) | ((cy - wy0) << 10) | (cz << 20));
                    if (pn < g.path_cap) path[pn] = (cz * g.Y + cy) * g.X + cx;
                    if (dir >= 4) via += 1;
                    else if (dir < 2) wl += abs(g.xc[cx] - g.xc[px]);
                    else wl += abs(g.yc[cy] - g.yc[py]);
                }
                pn += 1; pend += 1;
                cx = __shfl_sync(0xFFFFFFFFu, px, dir);
                cy = __shfl_sync(0xFFFFFFFFu, py, dir);
                cz = __shfl_sync(0xFFFFFFFFu, pz, dir);
                last = dir;
            }
            if (!fail) {
                if (first && pend + 1 > WIN_TGT_CAP) flush();
                if (lane == 0) {
                    if (first) s_tloc[pend] = (uint32_t)((cx - wx0) | ((cy - wy0) << 10) | (cz << 20));
                    if (pn < g.path_cap) path[pn] = (cz * g.Y + cy) * g.X + cx;
                }
                pn += 1;
                if (first) pend += 1;
            }
            flush();
            // The walk is over: only now do its cells join the tree on chip (distance 0, tree bit, dirty lines) -- a cell
            // zeroed under the walk could pass the predecessor test of a later cell.  They are read back from the path
            // record, which must therefore hold the whole net (an overflow fails the step with XR_E_CAPACITY).  The
            // cell the walk ended on is a source already; on the first connection it gets its tree bit here.
            const bool over = pn > g.path_cap;
            for (int k = pn0 + lane; k < (over ? g.path_cap : pn); k += 32) {
                const int ci = path[k];
                const int x = ci % g.X - wx0, y = (ci / g.X) % g.Y - wy0, z = ci / (g.X * g.Y);
                uint32_t *pq = win_cell_ptr<C>(cluster, c, H, x, y, z);
                *pq = (*pq & ~WMASK) | (CF_TREE << 28);
                win_mark<C>(cluster, c, x, y, z);
            }
            fail |= over;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                wl += __shfl_xor_sync(0xFFFFFFFFu, wl, off);
                via += __shfl_xor_sync(0xFFFFFFFFu, via, off);
            }
            if (lane == 0) {
                d.wlvia[2 * env] += wl; d.wlvia[2 * env + 1] += via;
                d.path_n[env] = pn;
                if (cn < g.conn_cap) {
                    d.conn_cost[(size_t)env * g.conn_cap + cn] = B;
                    d.conn_off[(size_t)env * (g.conn_cap + 1) + cn + 1] = pn;
                }
                d.conn_n[env] = cn + 1;
                d.envstat[8 * (size_t)env + 3] += 1;
                if (fail) d.flags[1] = over ? 4 : 3;
                s_flag[3] = fail ? 0 : 1;
            }
        }
        if (C > 1) cluster.sync(); else __syncthreads();
        const int go_on = (C > 1) ? *cluster.map_shared_rank(&s_flag[3], 0) : s_flag[3];
        // ---- pin bookkeeping, by every CTA on its own copy: a pin is connected once any of its access points is on
        // the tree (tree bit of the on-chip cell); rank 0 mirrors the flags to global memory for a hand-over
        for (int i = tid; i < n_ap; i += WIN_T) {
            const uint32_t q = s_apc[i];
            s_apon[i] = s_apconn[i] ? 1 : (((*win_cell_ptr<C>(cluster, c, H, q & 1023, (q >> 10) & 1023, q >> 20)) >> 31) & 1u);
        }
        __syncthreads();
        bool left = false;
        for (int i = tid; i < n_ap; i += WIN_T) {
            if (s_apconn[i]) continue;
            const unsigned pin = s_appin[i];
            bool on = false;
            for (int j = i; j >= 0 && s_appin[j] == pin && !on; j--) on = s_apon[j] != 0;
            for (int j = i + 1; j < n_ap && s_appin[j] == pin && !on; j++) on = s_apon[j] != 0;
            if (on) { s_apconn[i] = 1; if (rank == 0) d.ap_conn[aoff + i] = 1; }
            else left = true;
        }
        const int more = __syncthreads_or(left) && go_on;
#ifdef WIN_PHASE_SPLIT
        sp[5] += clock64() - tq0;           // target choice, exit test, backtrace, commit, pin bookkeeping
#endif
        if (!more) break;
        if (first) {
            // after the first connection only the path is the tree: the unused APs of the
            // source pin leave the source set, so the field is rebuilt from the tree.  With a
            // single source AP the source set only grew and the old field stays a valid bound.
            if (n_src_ap > 1) {
                for (int i = tid; i < c.Z * c.HH * c.WXp; i += WIN_T) {
                    const uint32_t v = c.cell[i];
                    const bool tree = ((v >> 28) & CF_TREE) != 0;
                    c.cell[i] = (v & ~WMASK) | (tree ? 0u : WINF);
                    if (tree) {
                        const int x = i % c.WXp, ly = (i / c.WXp) % c.HH, z = i / (c.WXp * c.HH);
                        if (x < WX && ly >= 1 && ly <= c.h) {
                            c.rowd[z * H + ly - 1] = 1; c.cold[z * WX + x] = 1; c.posd[(ly - 1) * WX + x] = 1;
                        }
                    }
                }
            }
            first = false;
        }
        if (C > 1) cluster.sync(); else __syncthreads();
    }
    // relaxation accounting (cells touched by the in-window sweeps)
#ifndef WIN_PHASE_RANK
#define WIN_PHASE_RANK 0
#endif
#ifndef WIN_PHASE_MINC
#define WIN_PHASE_MINC 1
#endif
    if (tid == 0 && rank == (C > WIN_PHASE_RANK ? WIN_PHASE_RANK : 0) && d.dbg) {
        atomicAdd(&d.dbg[0], (unsigned long long)n_iter); atomicAdd(&d.dbg[1], (unsigned long long)n_conn);
        atomicAdd(&d.dbg[2], (unsigned long long)cyc_relax); atomicAdd(&d.dbg[3], (unsigned long long)(clock64() - tk0));
        atomicAdd(&d.dbg[4], 1ull); atomicAdd(&d.dbg[5], (unsigned long long)(WX * WY));
        if (C == 8) { atomicAdd(&d.dbg[6], (unsigned long long)n_iter); atomicAdd(&d.dbg[7], (unsigned long long)cyc_relax);
                      atomicAdd(&d.dbg[15], (unsigned long long)n_conn); }
#ifdef WIN_PHASE_TIMING
#ifdef WIN_PHASE_SPLIT
        if (C >= WIN_PHASE_MINC) { for (int k = 0; k < 6; k++) atomicAdd(&d.dbg[8 + k], (unsigned long long)sp[k]); }
#elif defined(WIN_PHASE_CRIT)
        if (C >= WIN_PHASE_MINC) { for (int k = 0; k < 6; k++) atomicAdd(&d.dbg[8 + k], (unsigned long long)phc[k]); atomicAdd(&d.dbg[14], (unsigned long long)ph[6]); }
#else
        if (C >= WIN_PHASE_MINC) for (int k = 0; k < 7; k++) atomicAdd(&d.dbg[8 + k], (unsigned long long)ph[k]);
#endif
#endif
    }
    // relaxation accounting: cells actually touched by the dirty-line sweeps of this band
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) work += __shfl_xor_sync(0xFFFFFFFFu, work, off);
    if (lane == 0 && work)
        atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 7]), (unsigned long long)work);
    if (tid == 0 && rank == 0)
        atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 2]), (unsigned long long)(3 * n_iter));
    if (C > 1) cluster.sync();                             // keep peers' shared memory alive until all are done
}
