// xr_kernels_maze.cuh -- maze route of the selected net for every environment.
//
// Replaces the external TritonRoute maze search the reference reaches through
// ZMQ (baseline/baseline_utils.py:409-419).  Specification: DESIGN.md section 3;
// the CPU checker used by the tests restates the same specification independently.
//
// The search is a Bellman-Ford fixpoint computed with line sweeps (GAMER style):
//   k_sweep_xz : for every row (y,z) a forward and a backward min-plus scan along
//                x (warp shuffles over (W,D) function pairs), then an up/down via
//                relaxation across the Z layers of that y-slice through shared
//                memory.  One read and one (conditional) write of dist per cell.
//   k_sweep_y  : for every (z, 32-column slab) the same scan along y, tiled over
//                warps with the carries composed through shared memory.
//   k_control  : per environment, one warp: convergence test, target selection,
//                canonical backtrace (32 cells of a straight run per memory round
//                trip), commit, pin bookkeeping, next connection or finish.
// Work is bounded three ways, none of which changes a result:
//   dirty lines   a row is swept along x (and its via stacks relaxed) only if one of its cells changed since
//                 its last sweep, a (layer, 32-column slab) along y likewise; blocks of clean lines exit at once;
//   bounded stop  a connection's relaxation ends as soon as every distance written in a pump is >= the best
//                 target distance (everything below it is final); pending dirty lines carry over;
//   cap           a relaxation whose result exceeds the best target distance known so far is skipped and its
//                 line flagged "deferred"; deferred lines turn dirty again when the next connection starts.
// The fixpoint of the relaxations is unique, so the distance field equals
// Dijkstra's regardless of sweep order; path identity rests on the canonical
// backtrace rule shared with the oracle.
#pragma once
#include "xr_common.cuh"

// ---------------------------------------------------------------- route begin
// Freeze the per-cell cost flags of the net each environment routes this step
// and clear its distance field.  4 cells per thread, 16-byte stores.
// Only the full-grid path reads these two arrays: environments whose net goes to a window kernel (mode 1) skip the
// pass -- the window kernel derives its window's flags from cellinfo / apnet itself -- and get it lazily
// (handover = 1, followed by k_handover_seed) in the rare step in which their search escapes the window.
__global__ void __launch_bounds__(256) k_route_begin(Geo g, Dev d, int grp, int handover) {
    const int env = blockIdx.y;
    const int net = d.act[2 * env + 1];
    if (net == 0) return;
    if (handover) { if (d.mode[env] != 1 || d.phase[env] != 2) return; }
    else if (d.mode[env] == 1 || (grp >= 0 && d.grp[env] != grp)) return;  // grp >= 0: only this post-route group
    const size_t eoff = (size_t)env * g.cells_p;
    const int n4 = g.cells_p >> 2;
    const uint4 *ci4 = reinterpret_cast<const uint4 *>(d.cellinfo + eoff);
    const uint2 *an4 = reinterpret_cast<const uint2 *>(d.apnet + eoff);
    uint4 *d4 = reinterpret_cast<uint4 *>(d.dist + eoff);
    uint32_t *f4 = reinterpret_cast<uint32_t *>(d.cflag + eoff);
    const uint4 inf4 = make_uint4(XR_INF, XR_INF, XR_INF, XR_INF);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const uint4 ci = __ldg(ci4 + i);
        const uint2 an = __ldg(an4 + i);
        const uint32_t c[4] = {ci.x, ci.y, ci.z, ci.w};
        const uint32_t a[4] = {an.x & 0xFFFFu, an.x >> 16, an.y & 0xFFFFu, an.y >> 16};
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) packed |= cost_flags(c[k], a[k], (uint32_t)net) << (8 * k);
        f4[i] = packed;
        d4[i] = inf4;
    }
}

// Second half of the lazy prologue of a hand-over (after k_route_begin<handover>): on the first connection the sources
// are the access points of the source pin; later the tree is what the window kernel committed -- the cells of the
// path record (the flags just rebuilt count the net's own wires as foreign route shapes, but only on tree cells,
// whose flags no relaxation and no backtrace ever reads: they are entered at distance 0 or not at all).
// grid N, block 64.
__global__ void k_handover_seed(Geo g, Dev d) {
    const int env = blockIdx.x;
    const int net = d.act[2 * env + 1];
    if (net == 0 || d.mode[env] != 1 || d.phase[env] != 2) return;
    const size_t eoff = (size_t)env * g.cells_p;
    __syncthreads();                                      // (everyone has read the phase)
    if (threadIdx.x == 0) { d.phase[env] = 1; atomicAdd(&d.flags[0], 1); atomicSub(&d.flags[3], 1); }   // armed: the pumps take it from here
    if (d.first[env]) {
        const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
        const unsigned srcpin = d.net_srcpin[(size_t)env * (g.max_nets + 1) + net];
        const size_t aoff = (size_t)env * g.max_aps;
        for (int i = ns[net] + threadIdx.x; i < ns[net + 1]; i += blockDim.x)
            if (d.ap_pin[aoff + i] == srcpin) d.dist[eoff + d.ap_cellp[aoff + i]] = 0;
        return;
    }
    const int *path = d.path + (size_t)env * g.path_cap;
    const int pn = d.path_n[env] < g.path_cap ? d.path_n[env] : g.path_cap;
    for (int k = threadIdx.x; k < pn; k += blockDim.x) {
        const int ci = path[k];
        const int x = ci % g.X, y = (ci / g.X) % g.Y, z = ci / (g.X * g.Y);
        d.cflag[eoff + ((size_t)z * g.Y + y) * g.Xp + x] = CF_TREE;       // (its cost flags are never read, see above)
    }
}

// Per-environment step prologue: mark the chosen net routed, arm the route state
// machine and seed the sources (all APs of the static source pin).
__global__ void k_seed(Geo g, Dev d, int grp) {
    const int env = blockIdx.x;
    if (grp >= 0 && d.grp[env] != grp) return;
    const int raw = d.act[2 * env], net = d.act[2 * env + 1];
    if (threadIdx.x == 0) {
        d.obs_do[env] = (raw != 0);
        d.fin[env] = 0;
        if (raw >= 1) d.routed[(size_t)env * (g.max_nets + 1) + raw] = 1;
        if (raw == -1) d.done[env] = 1;
        d.path_n[env] = 0; d.conn_n[env] = 0;
        d.conn_off[(size_t)env * (g.conn_cap + 1)] = 0;
    }
    if (net == 0) return;
    if (d.mode[env] == 2) {                               // frontier engine: it keeps its own field and source set
        if (threadIdx.x == 0) { d.changed[env] = 0; d.reinit[env] = 0; d.first[env] = 1; d.phase[env] = 0; }
        return;
    }
    // full-grid bookkeeping: the first pump of a net (or of a hand-over from a window kernel) sweeps everything
    {
        const int S = g.Xp / 32;
        for (int i = threadIdx.x; i < g.Y; i += blockDim.x) { d.g_rowd[(size_t)env * g.Y + i] = 0; d.g_rowf[(size_t)env * g.Y + i] = 0; }
        for (int i = threadIdx.x; i < g.Z * S; i += blockDim.x) { d.g_slabd[(size_t)env * g.Z * S + i] = 0; d.g_slabf[(size_t)env * g.Z * S + i] = 0; }
        if (threadIdx.x == 0) { d.g_all[env] = 1; d.g_cap[env] = XR_INF; d.g_gmin[env] = 0xFFFFFFFFu; }
    }
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const int s = ns[net], t = ns[net + 1];
    const unsigned srcpin = d.net_srcpin[(size_t)env * (g.max_nets + 1) + net];
    const size_t aoff = (size_t)env * g.max_aps;
    for (int i = s + threadIdx.x; i < t; i += blockDim.x) {
        const bool src = d.ap_pin[aoff + i] == srcpin;
        d.ap_conn[aoff + i] = src;
        if (src) {
            const size_t c = (size_t)env * g.cells_p + d.ap_cellp[aoff + i];
            d.dist[c] = 0;
        }
    }
    if (threadIdx.x == 0) {
        d.changed[env] = 1; d.reinit[env] = 0; d.first[env] = 1;
        if (d.mode[env] == 0) {               // global path: arm the state machine now; the
            d.phase[env] = 1;                 // window kernel arms it only if it has to fall back
            atomicAdd(&d.flags[0], 1);
        } else d.phase[env] = 0;
    }
}

// ------------------------------------------------------------------ sweep x+z
// Weight of entering cell x of a row on layer z from the left (x-1 -> x) / right.
__device__ __forceinline__ uint32_t wx_from_left(const Geo &g, int z, int x, uint32_t f) {
    if (x < 1 || x >= g.X) return XR_INF;
    return wgt_x(g, z, g.uniform_x ? (uint32_t)g.dx : (uint32_t)(g.xc[x] - g.xc[x - 1]), f);
}
__device__ __forceinline__ uint32_t wx_from_right(const Geo &g, int z, int x, uint32_t f) {
    if (x >= g.X - 1) return XR_INF;
    return wgt_x(g, z, g.uniform_x ? (uint32_t)g.dx : (uint32_t)(g.xc[x + 1] - g.xc[x]), f);
}

// grid (Y, N), block (32, Z).  Warp z owns row (y, z); lane l owns CPL consecutive
// cells starting at x0 = l*CPL.  Flags are kept 4 bits per cell in registers.
#ifndef XZ_ROWS
#define XZ_ROWS 1                    // rows per block (more: fewer launches of clean-row blocks, but a flood's rows serialise)
#endif
template <int CPL>
__device__ __forceinline__ void sweep_xz_row(const Geo &g, const Dev &d, int env, int y, bool re, uint32_t *sd, uint8_t *sf) {
    const int z = threadIdx.y, lane = threadIdx.x;
    uint8_t *rowd = d.g_rowd + (size_t)env * g.Y + y;
    if (!re && d.g_all[env] == 0 && *rowd == 0) return;          // clean row (uniform over the block)
    __syncthreads();                                             // (also: the previous row's via stage is done with sd/sf)
    if (z == 0 && lane == 0) *rowd = 0;
    const uint32_t cap = d.g_cap[env];
    bool skipped = false;
    uint32_t gm = 0xFFFFFFFFu;
    const size_t rowoff = (size_t)env * g.cells_p + ((size_t)z * g.Y + y) * g.Xp;
    const int x0 = lane * CPL;
    constexpr int NF = (CPL + 7) / 8;
    uint32_t dd[CPL], fpk[NF];
#pragma unroll
    for (int i = 0; i < NF; i++) fpk[i] = 0;
    const bool have = x0 < g.Xp;
    // ---- load the row (vectorised, coalesced)
    if (have) {
        if (CPL >= 4) {
            const uint4 *p = reinterpret_cast<const uint4 *>(d.dist + rowoff + x0);
            const uint32_t *q = reinterpret_cast<const uint32_t *>(d.cflag + rowoff + x0);
#pragma unroll
            for (int v = 0; v < CPL / 4; v++) {
                const uint4 t = p[v];
                dd[4 * v] = t.x; dd[4 * v + 1] = t.y; dd[4 * v + 2] = t.z; dd[4 * v + 3] = t.w;
                const uint32_t fw = q[v];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    fpk[(4 * v + k) >> 3] |= ((fw >> (8 * k)) & 0xFu) << (4 * ((4 * v + k) & 7));
            }
        } else {
#pragma unroll
            for (int i = 0; i < CPL; i++) {
                dd[i] = d.dist[rowoff + x0 + i];
                fpk[i >> 3] |= ((uint32_t)d.cflag[rowoff + x0 + i] & 0xFu) << (4 * (i & 7));
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < CPL; i++) dd[i] = XR_INF;
    }
#define FLG(i) ((fpk[(i) >> 3] >> (4 * ((i) & 7))) & 0xFu)
    uint32_t chg = 0;                    // bit i: cell i must be written back
#pragma unroll
    for (int i = 0; i < CPL; i++) {
        if (re) { dd[i] = (FLG(i) & CF_TREE) ? 0u : XR_INF; chg |= 1u << i; }
        if (x0 + i >= g.X) dd[i] = XR_INF;
    }
    // ---- forward scan: cell i is f_i(t) = min(t + w_i, d_i); compose left to right
    {
        uint32_t W = 0, D = XR_INF;
#pragma unroll
        for (int i = 0; i < CPL; i++) {
            const uint32_t w = wx_from_left(g, z, x0 + i, FLG(i));
            D = xr_min(D + w, dd[i]); W = xr_min(W + w, XR_INF);
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t Wo = __shfl_up_sync(0xFFFFFFFFu, W, off);
            const uint32_t Do = __shfl_up_sync(0xFFFFFFFFu, D, off);
            if (lane >= off) { D = xr_min(Do + W, D); W = xr_min(Wo + W, XR_INF); }
        }
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, D, 1);
        if (lane == 0) t = XR_INF;
#pragma unroll
        for (int i = 0; i < CPL; i++) {
            t = xr_min(t + wx_from_left(g, z, x0 + i, FLG(i)), dd[i]);
            if (t < dd[i]) {
                if (t <= cap) { chg |= 1u << i; dd[i] = t; gm = xr_min(gm, t); } else skipped = true;
            }
        }
    }
    // ---- backward scan
    {
        uint32_t W = 0, D = XR_INF;
#pragma unroll
        for (int i = CPL - 1; i >= 0; i--) {
            const uint32_t w = wx_from_right(g, z, x0 + i, FLG(i));
            D = xr_min(D + w, dd[i]); W = xr_min(W + w, XR_INF);
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t Wo = __shfl_down_sync(0xFFFFFFFFu, W, off);
            const uint32_t Do = __shfl_down_sync(0xFFFFFFFFu, D, off);
            if (lane + off < 32) { D = xr_min(Do + W, D); W = xr_min(Wo + W, XR_INF); }
        }
        uint32_t t = __shfl_down_sync(0xFFFFFFFFu, D, 1);
        if (lane == 31) t = XR_INF;
#pragma unroll
        for (int i = CPL - 1; i >= 0; i--) {
            t = xr_min(t + wx_from_right(g, z, x0 + i, FLG(i)), dd[i]);
            if (t < dd[i]) {
                if (t <= cap) { chg |= 1u << i; dd[i] = t; gm = xr_min(gm, t); } else skipped = true;
            }
        }
    }
    // ---- stage the slice for the via relaxation (bit 7 of the flag byte = "write back")
    if (have) {
#pragma unroll
        for (int i = 0; i < CPL; i++) {
            const int x = x0 + i;
            sd[z * g.Xp + x] = dd[i];
            sf[z * g.Xp + x] = (uint8_t)(FLG(i) | (((chg >> i) & 1u) ? 0x80u : 0u));
        }
    }
#undef FLG
    __syncthreads();
    const int tid = threadIdx.y * 32 + threadIdx.x, nthr = blockDim.y * 32;
    bool any = false, via_any = false;
    const int S = g.Xp / 32;
    for (int x = tid; x < g.X; x += nthr) {
        uint32_t v[XR_ZMAX], fz[XR_ZMAX];
#pragma unroll
        for (int k = 0; k < XR_ZMAX; k++) if (k < g.Z) { v[k] = sd[k * g.Xp + x]; fz[k] = sf[k * g.Xp + x]; }
        uint32_t t = v[0];
#pragma unroll
        for (int k = 1; k < XR_ZMAX; k++) if (k < g.Z) {
            t = xr_min(t + wgt_v(g, k - 1, k, fz[k]), v[k]);
            if (t < v[k]) {
                if (t <= cap) { v[k] = t; fz[k] |= 0x80u; gm = xr_min(gm, t); via_any = true; } else skipped = true;
            }
        }
#pragma unroll
        for (int k = XR_ZMAX - 2; k >= 0; k--) if (k < g.Z - 1) {
            t = xr_min(t + wgt_v(g, k, k, fz[k]), v[k]);
            if (t < v[k]) {
                if (t <= cap) { v[k] = t; fz[k] |= 0x80u; gm = xr_min(gm, t); via_any = true; } else skipped = true;
            }
        }
#pragma unroll
        for (int k = 0; k < XR_ZMAX; k++) if (k < g.Z) {
            if (fz[k] & 0x80u) {
                d.dist[(size_t)env * g.cells_p + ((size_t)k * g.Y + y) * g.Xp + x] = v[k];
                d.g_slabd[((size_t)env * g.Z + k) * S + (x >> 5)] = 1;          // this column has to be swept along y
                any = true;
            }
        }
    }
    if (via_any) *rowd = 1;                                 // a via move lowered a cell: the row is swept again
    if (skipped) d.g_rowf[(size_t)env * g.Y + y] = 1;
    gm = __reduce_min_sync(0xFFFFFFFFu, gm);
    if (lane == 0 && gm != 0xFFFFFFFFu) atomicMin(&d.g_gmin[env], gm);
    if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 7]), (unsigned long long)g.Z * g.X);
    if (__any_sync(0xFFFFFFFFu, any) && lane == 0) d.changed[env] = 1;
}
// grid (ceil(Y / XZ_ROWS), N), block (32, Z)
template <int CPL>
__global__ void __launch_bounds__(32 * XR_ZMAX) k_sweep_xz(Geo g, Dev d) {
    const int env = blockIdx.y;
    if (d.phase[env] != 1) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *sd = reinterpret_cast<uint32_t *>(smem_raw);                       // [Z][Xp]
    uint8_t *sf = smem_raw + (size_t)g.Z * g.Xp * 4;                              // [Z][Xp]
    const bool re = d.reinit[env] != 0;
    const int y1 = min(g.Y, (int)(blockIdx.x + 1) * XZ_ROWS);
    for (int y = blockIdx.x * XZ_ROWS; y < y1; y++) sweep_xz_row<CPL>(g, d, env, y, re, sd, sf);
}

// -------------------------------------------------------------------- sweep y
// grid (Xp/32, Z, N), block (32, NW), NW = ceil(Y / TH) <= 16.  Warp w owns rows
// [w*TH, w*TH+TH) of a 32-column slab; lane = column.
template <int TH>
__global__ void __launch_bounds__(32 * 16) k_sweep_y(Geo g, Dev d) {
    const int env = blockIdx.z;
    if (d.phase[env] != 1) return;
    __shared__ uint32_t sW[2][16][32], sD[2][16][32];
    const int z = blockIdx.y, lane = threadIdx.x, w = threadIdx.y, nw = blockDim.y;
    uint8_t *slabd = d.g_slabd + ((size_t)env * g.Z + z) * (g.Xp / 32) + blockIdx.x;
    if (d.g_all[env] == 0 && *slabd == 0) return;                // clean slab (uniform over the block)
    __syncthreads();
    if (w == 0 && lane == 0) *slabd = 0;
    const uint32_t cap = d.g_cap[env];
    bool skipped = false;
    uint32_t gm = 0xFFFFFFFFu;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = w * TH;
    const bool colok = x < g.X;
    const size_t base = (size_t)env * g.cells_p + (size_t)z * g.Y * g.Xp + x;
    uint32_t dd[TH];
    uint32_t fpk[(TH + 7) / 8];          // 4 bits of flags per cell
#pragma unroll
    for (int i = 0; i < (TH + 7) / 8; i++) fpk[i] = 0;
#pragma unroll
    for (int i = 0; i < TH; i++) {
        const int y = y0 + i;
        uint32_t v = XR_INF, f = 0;
        if (colok && y < g.Y) { v = d.dist[base + (size_t)y * g.Xp]; f = d.cflag[base + (size_t)y * g.Xp] & 7u; }
        dd[i] = v; fpk[i >> 3] |= f << (4 * (i & 7));
    }
    unsigned long long chg = 0ull;
    // forward: entering row y from y-1
    {
        uint32_t W = 0, D = XR_INF;
#pragma unroll
        for (int i = 0; i < TH; i++) {
            const int y = y0 + i;
            const uint32_t f = (fpk[i >> 3] >> (4 * (i & 7))) & 7u;
            uint32_t wv = XR_INF;
            if (y >= 1 && y < g.Y) wv = wgt_y(g, z, g.uniform_y ? (uint32_t)g.dy : (uint32_t)(g.yc[y] - g.yc[y - 1]), f);
            D = xr_min(D + wv, dd[i]); W = xr_min(W + wv, XR_INF);
        }
        sW[0][w][lane] = W; sD[0][w][lane] = D;
        __syncthreads();
        uint32_t t = XR_INF;
        for (int j = 0; j < w; j++) t = xr_min(t + sW[0][j][lane], sD[0][j][lane]);
#pragma unroll
        for (int i = 0; i < TH; i++) {
            const int y = y0 + i;
            const uint32_t f = (fpk[i >> 3] >> (4 * (i & 7))) & 7u;
            uint32_t wv = XR_INF;
            if (y >= 1 && y < g.Y) wv = wgt_y(g, z, g.uniform_y ? (uint32_t)g.dy : (uint32_t)(g.yc[y] - g.yc[y - 1]), f);
            t = xr_min(t + wv, dd[i]);
            if (t < dd[i]) {
                if (t <= cap) { chg |= 1ull << i; dd[i] = t; gm = xr_min(gm, t); } else skipped = true;
            }
        }
    }
    // backward: entering row y from y+1
    {
        uint32_t W = 0, D = XR_INF;
#pragma unroll
        for (int i = TH - 1; i >= 0; i--) {
            const int y = y0 + i;
            const uint32_t f = (fpk[i >> 3] >> (4 * (i & 7))) & 7u;
            uint32_t wv = XR_INF;
            if (y < g.Y - 1) wv = wgt_y(g, z, g.uniform_y ? (uint32_t)g.dy : (uint32_t)(g.yc[y + 1] - g.yc[y]), f);
            D = xr_min(D + wv, dd[i]); W = xr_min(W + wv, XR_INF);
        }
        sW[1][w][lane] = W; sD[1][w][lane] = D;
        __syncthreads();
        uint32_t t = XR_INF;
        for (int j = nw - 1; j > w; j--) t = xr_min(t + sW[1][j][lane], sD[1][j][lane]);
#pragma unroll
        for (int i = TH - 1; i >= 0; i--) {
            const int y = y0 + i;
            const uint32_t f = (fpk[i >> 3] >> (4 * (i & 7))) & 7u;
            uint32_t wv = XR_INF;
            if (y < g.Y - 1) wv = wgt_y(g, z, g.uniform_y ? (uint32_t)g.dy : (uint32_t)(g.yc[y + 1] - g.yc[y]), f);
            t = xr_min(t + wv, dd[i]);
            if (t < dd[i]) {
                if (t <= cap) { chg |= 1ull << i; dd[i] = t; gm = xr_min(gm, t); } else skipped = true;
            }
        }
    }
    if (colok) {
#pragma unroll
        for (int i = 0; i < TH; i++)
            if ((chg >> i) & 1ull) {
                d.dist[base + (size_t)(y0 + i) * g.Xp] = dd[i];
                d.g_rowd[(size_t)env * g.Y + y0 + i] = 1;                      // this row has to be swept along x
            }
    }
    if (skipped) d.g_slabf[((size_t)env * g.Z + z) * (g.Xp / 32) + blockIdx.x] = 1;
    gm = __reduce_min_sync(0xFFFFFFFFu, gm);
    if (lane == 0 && gm != 0xFFFFFFFFu) atomicMin(&d.g_gmin[env], gm);
    if (w == 0 && lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 7]), 32ull * g.Y);
    if (__any_sync(0xFFFFFFFFu, chg != 0ull) && lane == 0) d.changed[env] = 1;
}

// -------------------------------------------------------------------- control
// Commit one path cell: occupancy, obstacle channel source, tree mark, new source.
// zero_dist = false: the caller is still walking the distance field (full-grid backtrace) and turns the path cells into
// sources only after the walk -- a cell that became 0 under the walk could pass the predecessor test of a later cell
// (found by tools/fuzz_parity.py on a non-uniform grid: x step == 3 * y step next to the source).
template <bool zero_dist = true>
__device__ __forceinline__ void commit_cell(const Geo &g, const Dev &d, int env, int net, int x, int y, int z) {
    const size_t c = (size_t)env * g.cells_p + ((size_t)z * g.Y + y) * g.Xp + x;
    uint32_t ci = d.cellinfo[c];
    {   // congestion counts maintained by the commits (XrConfig.metrics_mode 0; k_metrics recomputes them by a scan)
        const uint32_t us = (ci & CI_USAGE_MASK) >> CI_USAGE_SHIFT;
        unsigned long long *mc = reinterpret_cast<unsigned long long *>(d.minc + 4 * (size_t)env);
        if (us == 0u) {
            if (ci & CI_BLOCK) atomicAdd(mc + 0, 1ull);
            if ((ci & CI_ISAP) && d.apnet[c] != (uint16_t)net) atomicAdd(mc + 1, 1ull);
        } else if (us == 1u) {
            if (!((ci & CI_ISAP) && d.apnet[c] != (ci & CI_OWNER_MASK))) atomicAdd(mc + 1, 1ull);
            atomicAdd(mc + 2, 1ull);
        } else if (us < 255u) atomicAdd(mc + 2, 1ull);
    }
    if (((ci & CI_USAGE_MASK) >> CI_USAGE_SHIFT) < 255u) ci += 1u << CI_USAGE_SHIFT;
    if ((ci & CI_OWNER_MASK) == 0u) ci |= (uint32_t)net;
    d.cellinfo[c] = ci;
    atomicOr(d.dist64 + c, FRW_RS);                         // frontier cell word: a wire covers the cell
    const size_t oo = ((size_t)x * g.Y + y) * g.Z + z;
    d.obst_obs[(size_t)env * g.cells_o + oo] = 1;
    d.obs[(size_t)env * g.obs_stride + oo] = 1.f;        // channel 0 of the observation, updated in place
    d.cflag[c] |= CF_TREE;
    if (zero_dist) d.dist[c] = 0;
    d.g_rowd[(size_t)env * g.Y + y] = 1;                    // full-grid path: the new source's lines are dirty
    d.g_slabd[((size_t)env * g.Z + z) * (g.Xp / 32) + (x >> 5)] = 1;
}

// grid N, block 32 (one warp per environment).
__global__ void __launch_bounds__(32) k_control(Geo g, Dev d) {
    const int env = blockIdx.x, lane = threadIdx.x;
    if (d.phase[env] != 1) return;
    if (lane == 0) d.envstat[8 * (size_t)env + 2] += 2;  // one more pump = 2 relaxation passes (over the dirty lines)
    const int net = d.act[2 * env + 1];
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const int s = ns[net], t = ns[net + 1];
    const size_t aoff = (size_t)env * g.max_aps;
    const size_t eoff = (size_t)env * g.cells_p;
    const bool first = d.first[env] != 0;
    if (d.changed[env] != 0) {                           // something was written in this pump
        // bounded stop: once every distance written in a pump is >= the best target distance, all cells below
        // it are final; otherwise tighten the cap and pump again
        uint32_t bcur = 0xFFFFFFFFu;
        for (int i = s + lane; i < t; i += 32)
            if (!d.ap_conn[aoff + i]) bcur = xr_min(bcur, d.dist[eoff + d.ap_cellp[aoff + i]]);
        bcur = __reduce_min_sync(0xFFFFFFFFu, bcur);
        const uint32_t gm = d.g_gmin[env];
        const bool was_reinit = d.reinit[env] != 0;
        __syncwarp();
        if (lane == 0) {
            d.changed[env] = 0; d.reinit[env] = 0; d.g_all[env] = 0; d.g_gmin[env] = 0xFFFFFFFFu;
            if (bcur < d.g_cap[env]) d.g_cap[env] = bcur;
        }
        if (was_reinit || !(bcur < XR_INF && gm >= bcur)) return;
    }
    // ---- target = argmin (dist, cell index) over the APs of unconnected pins
    unsigned long long best = ~0ull;
    for (int i = s + lane; i < t; i += 32) {
        if (d.ap_conn[aoff + i]) continue;
        const int cp = d.ap_cellp[aoff + i];
        const unsigned long long key = ((unsigned long long)d.dist[eoff + cp] << 32) | (unsigned)cp;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, best, off);
        best = o < best ? o : best;
    }
    const uint32_t tdist = (uint32_t)(best >> 32);
    if (best == ~0ull || tdist >= XR_INF) {              // nothing to connect / unreachable
        if (lane == 0) {
            if (best != ~0ull) d.flags[1] = 1;
            d.phase[env] = 0; atomicSub(&d.flags[0], 1);
        }
        return;
    }
    int cp = (int)(best & 0xFFFFFFFFu);
    int cx = cp % g.Xp, cy = (cp / g.Xp) % g.Y, cz = cp / (g.Xp * g.Y);
    // ---- canonical backtrace + commit
    int pn = d.path_n[env];
    const int pn0 = pn;
    const int cn = d.conn_n[env];
    int *path = d.path + (size_t)env * g.path_cap;
    long long wl = 0, via = 0;
    int last = -1;
    bool fail = false;
    for (;;) {
        __syncwarp();
        const uint32_t dc = d.dist[eoff + ((size_t)cz * g.Y + cy) * g.Xp + cx];
        if (dc == 0) break;
        int run = 0;
        if (last >= 0) {
            int ddx, ddy, ddz; dir_delta(last, ddx, ddy, ddz);
            const int ax = cx - lane * ddx, ay = cy - lane * ddy, az = cz - lane * ddz;
            const int bx = ax - ddx, by = ay - ddy, bz = az - ddz;
            bool ok = ax >= 0 && ax < g.X && ay >= 0 && ay < g.Y && az >= 0 && az < g.Z &&
                      bx >= 0 && bx < g.X && by >= 0 && by < g.Y && bz >= 0 && bz < g.Z;
            if (ok) {
                const size_t ia = eoff + ((size_t)az * g.Y + ay) * g.Xp + ax;
                const size_t ib = eoff + ((size_t)bz * g.Y + by) * g.Xp + bx;
                const uint32_t da = d.dist[ia], db = d.dist[ib];
                ok = da != 0 && db < XR_INF && db + move_w(g, bx, by, bz, last, d.cflag[ia]) == da;
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            run = (m == 0xFFFFFFFFu) ? 32 : (__ffs(~m) - 1);
            if (run > 0) {
                if (lane < run) {
                    commit_cell<false>(g, d, env, net, ax, ay, az);
                    if (pn + lane < g.path_cap) path[pn + lane] = (az * g.Y + ay) * g.X + ax;
                    if (last >= 4) via += 1;
                    else if (last < 2) wl += abs(g.xc[ax] - g.xc[bx]);
                    else wl += abs(g.yc[ay] - g.yc[by]);
                }
                pn += run;
                cx -= run * ddx; cy -= run * ddy; cz -= run * ddz;
                continue;
            }
        }
        // no straight continuation: first valid of +x,-x,+y,-y,+z,-z
        {
            bool ok = false;
            int px = 0, py = 0, pz = 0;
            if (lane < 6) {
                int ddx, ddy, ddz; dir_delta(lane, ddx, ddy, ddz);
                px = cx - ddx; py = cy - ddy; pz = cz - ddz;
                if (px >= 0 && px < g.X && py >= 0 && py < g.Y && pz >= 0 && pz < g.Z) {
                    const uint32_t dp = d.dist[eoff + ((size_t)pz * g.Y + py) * g.Xp + px];
                    const uint32_t fc = d.cflag[eoff + ((size_t)cz * g.Y + cy) * g.Xp + cx];
                    ok = dp < XR_INF && dp + move_w(g, px, py, pz, lane, fc) == dc;
                }
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
            if (m == 0u) { fail = true; break; }
            const int dir = __ffs(m) - 1;
            if (lane == dir) {
                commit_cell<false>(g, d, env, net, cx, cy, cz);
                if (pn < g.path_cap) path[pn] = (cz * g.Y + cy) * g.X + cx;
                if (dir >= 4) via += 1;
                else if (dir < 2) wl += abs(g.xc[cx] - g.xc[px]);
                else wl += abs(g.yc[cy] - g.yc[py]);
            }
            pn += 1;
            cx = __shfl_sync(0xFFFFFFFFu, px, dir);
            cy = __shfl_sync(0xFFFFFFFFu, py, dir);
            cz = __shfl_sync(0xFFFFFFFFu, pz, dir);
            last = dir;
        }
    }
    // final cell (already a source): part of the recorded path; committed only on the
    // first connection, when it is a source-pin AP that is not on the tree yet
    if (!fail) {
        if (lane == 0) {
            if (first) commit_cell<false>(g, d, env, net, cx, cy, cz);
            if (pn < g.path_cap) path[pn] = (cz * g.Y + cy) * g.X + cx;
        }
        pn += 1;
    }
    // the walk is over: its cells become sources of the next connection (read back from the path record, which must
    // therefore hold the whole net: an overflow is an error of the step, XR_E_CAPACITY)
    __syncwarp();
    const bool over = pn > g.path_cap;
    for (int k = pn0 + lane; k < (over ? g.path_cap : pn); k += 32) {
        const int ci = path[k];
        const int x = ci % g.X, y = (ci / g.X) % g.Y, z = ci / (g.X * g.Y);
        d.dist[eoff + ((size_t)z * g.Y + y) * g.Xp + x] = 0;
    }
    fail |= over;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        wl += __shfl_xor_sync(0xFFFFFFFFu, wl, off);
        via += __shfl_xor_sync(0xFFFFFFFFu, via, off);
    }
    __syncwarp();
    __threadfence_block();
    // ---- pin bookkeeping: a pin is connected once any of its APs is on the tree
    // (CF_TREE is only ever set by commit_cell; the source pin is connected from
    // the start, its unused APs drop out of the source set at the reinit sweep).
    for (int i = s + lane; i < t; i += 32) {
        if (d.ap_conn[aoff + i]) continue;
        const unsigned pin = d.ap_pin[aoff + i];
        bool on = false;
        for (int j = i; j >= s && d.ap_pin[aoff + j] == pin && !on; j--)
            on = (d.cflag[eoff + d.ap_cellp[aoff + j]] & CF_TREE) != 0;
        for (int j = i + 1; j < t && d.ap_pin[aoff + j] == pin && !on; j++)
            on = (d.cflag[eoff + d.ap_cellp[aoff + j]] & CF_TREE) != 0;
        if (on) d.ap_conn[aoff + i] = 2;                 // 2 = newly connected (visible next round)
    }
    __syncwarp();
    bool left = false;
    for (int i = s + lane; i < t; i += 32) {
        uint8_t v = d.ap_conn[aoff + i];
        if (v == 2) { d.ap_conn[aoff + i] = 1; v = 1; }
        left |= (v == 0);
    }
    left = __any_sync(0xFFFFFFFFu, left);
    if (left && !fail) {
        // next connection: every deferred line is dirty again; its cap is the best distance the field left by the
        // previous connections already shows on a target (an upper bound, the tree only grew) -- none after a
        // re-initialisation, whose field starts from scratch
        const int S = g.Xp / 32;
        for (int i = lane; i < g.Y; i += 32)
            if (d.g_rowf[(size_t)env * g.Y + i]) { d.g_rowf[(size_t)env * g.Y + i] = 0; d.g_rowd[(size_t)env * g.Y + i] = 1; }
        for (int i = lane; i < g.Z * S; i += 32)
            if (d.g_slabf[(size_t)env * g.Z * S + i]) { d.g_slabf[(size_t)env * g.Z * S + i] = 0; d.g_slabd[(size_t)env * g.Z * S + i] = 1; }
        uint32_t b0 = XR_INF;
        if (!first)
            for (int i = s + lane; i < t; i += 32)
                if (!d.ap_conn[aoff + i]) b0 = xr_min(b0, d.dist[eoff + d.ap_cellp[aoff + i]]);
        b0 = __reduce_min_sync(0xFFFFFFFFu, b0);
        if (lane == 0) d.g_cap[env] = b0 < XR_INF ? b0 : XR_INF;
    }
    if (lane == 0) {
        d.wlvia[2 * env] += wl; d.wlvia[2 * env + 1] += via;
        d.path_n[env] = pn;
        if (cn < g.conn_cap) {
            d.conn_cost[(size_t)env * g.conn_cap + cn] = tdist;
            d.conn_off[(size_t)env * (g.conn_cap + 1) + cn + 1] = pn;
        }
        d.conn_n[env] = cn + 1;
        d.envstat[8 * (size_t)env + 3] += 1;
        if (fail) d.flags[1] = over ? 4 : 2;
        if (left && !fail) {
            d.changed[env] = 1;
            d.reinit[env] = first ? 1 : 0;
            d.first[env] = 0;
            d.g_all[env] = first ? 1 : 0;                // the re-initialisation pass touches every line
            d.g_gmin[env] = 0xFFFFFFFFu;
        } else {
            d.phase[env] = 0;
            atomicSub(&d.flags[0], 1);
        }
    }
}
