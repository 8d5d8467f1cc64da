// xr_kernels_win2.cuh -- window-resident maze route, dual cyclic layout (many-pin nets).
//
// The band kernel (xr_kernels_win.cuh) gives every CTA of the cluster a band of rows.  A
// many-pin net spends most of its time in *later* connections, each of which lowers the
// distance field in a ball around the newly committed branch: that ball covers two or three
// bands, whose CTAs become the critical path of every iteration while the others idle
// (measured: busiest band 47 k cycles per iteration, idlest 4.7 k).
//
// Here the window is kept twice, both times distributed *cyclically*, so that any compact
// region is spread evenly over the cluster:
//   layout A  rows    y = rank + C*ya  of every layer, x fastest  -> x sweeps and via stacks
//   layout B  columns x = rank + C*xb  of every layer, y fastest  -> y sweeps
// No line crosses a CTA, so there are no halo rows and no pull.  A sweep writes its own
// layout in place and *pushes* every lowered cell to the owner of that cell in the other
// layout with a remote shared-memory store (st.shared::cluster, fire and forget) together
// with the dirty flag of the line that has to be walked there; the cluster barrier between
// the y phase and the x phase (and the one closing the iteration) publishes the pushes.
// Both copies are therefore equal whenever a phase starts, and the fixpoint -- hence every
// path, cost and metric -- is the one of the band kernel, the full-grid sweeps and the oracle.
#pragma once
#include "xr_kernels_win.cuh"

#define WIN2_CELL_WORDS(Z, C, WX, WY) \
    ((Z) * (((WY) + (C) - 1) / (C)) * ((WX) | 1) + (Z) * (((WX) + (C) - 1) / (C)) * ((WY) | 1))
#define WIN2_AUX_WORDS(Z, C, WX, WY)                                                                          \
    (30 * (Z) + (WX) + 2 + (WY) + 2 + 64 + 6 * WIN_TGT_CAP +                                                   \
     ((Z) * (((WY) + (C) - 1) / (C)) + (Z) * (((WX) + (C) - 1) / (C)) + (((WY) + (C) - 1) / (C)) * (WX)) / 4 + 3 + \
     ((Z) * ((((WY) + (C) - 1) / (C)) > (((WX) + (C) - 1) / (C)) ? (((WY) + (C) - 1) / (C)) : (((WX) + (C) - 1) / (C)))) / 2 + 2 + 8)

struct Win2Ctx {
    uint32_t *A;         // [Z][HA][WXp]
    uint32_t *B;         // [Z][WB][WYp]
    uint32_t *wlut, *lutm, *pens, *lenx, *leny;
    uint8_t *rowd;       // [Z*HA] dirty rows of A
    uint8_t *cold;       // [Z*WB] dirty columns of B
    uint8_t *posd;       // [HA*WX] dirty via stacks of A
    uint16_t *list;
    int *cnt;
    int Z, WX, WXp, WY, WYp, HA, WB, ha, wb, rank;
    int uni_x, uni_y;
    uint32_t sA, sB, sRowd, sCold, sPosd;   // the same arrays as 32-bit shared-window addresses (for mapa / st.shared::cluster)
};

// shared-window address of `addr` in CTA `r` of the cluster, and fire-and-forget stores to it
__device__ __forceinline__ uint32_t win2_mapa(uint32_t addr, uint32_t r) {
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(addr), "r"(r));
    return o;
}
__device__ __forceinline__ void win2_st32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void win2_st8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <int C> struct Log2C { static constexpr int v = C == 1 ? 0 : C == 2 ? 1 : C == 4 ? 2 : C == 8 ? 3 : 4; };

__device__ int win2_compact(const Win2Ctx &c, uint8_t *flags, int n) {
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

// A lowered cell of layout A at (z, ya, x): flag its via stack here, push it to its column in layout B.
template <int C>
__device__ __forceinline__ void win2_push_from_A(cg::cluster_group &cluster, const Win2Ctx &c, int z, int ya, int x, uint32_t v) {
    const int r = x & (C - 1), xb = x >> Log2C<C>::v, y = c.rank + C * ya;
    const int col = z * c.WB + xb;
    if (C > 1) {
        win2_st32(win2_mapa(c.sB + 4u * (uint32_t)(col * c.WYp + y), r), v);
        win2_st8(win2_mapa(c.sCold + (uint32_t)col, r), 1u);
    } else { c.B[(size_t)col * c.WYp + y] = v; c.cold[col] = 1; }
}
// A lowered cell of layout B at (z, xb, y): push it to its row in layout A and flag the row and the stack there.
template <int C>
__device__ __forceinline__ void win2_push_from_B(cg::cluster_group &cluster, const Win2Ctx &c, int z, int xb, int y, uint32_t v) {
    const int r = y & (C - 1), ya = y >> Log2C<C>::v, x = c.rank + C * xb;
    const int row = z * c.HA + ya;
    if (C > 1) {
        win2_st32(win2_mapa(c.sA + 4u * (uint32_t)(row * c.WXp + x), r), v);
        win2_st8(win2_mapa(c.sRowd + (uint32_t)row, r), 1u);
        win2_st8(win2_mapa(c.sPosd + (uint32_t)(ya * c.WX + x), r), 1u);
    } else { c.A[(size_t)row * c.WXp + x] = v; c.rowd[row] = 1; c.posd[ya * c.WX + x] = 1; }
}

// Thread-per-line walk in direction DIR over a line of n cells with stride 1.
// AXIS 0: row (z, own = ya) of layout A; AXIS 1: column (z, own = xb) of layout B.
template <int C, int AXIS, int DIR, bool UNI>
__device__ __forceinline__ uint32_t win2_walk(cg::cluster_group &cluster, const Win2Ctx &c, uint32_t *__restrict__ p, int n,
                                              uint32_t lutreg, uint32_t pen, const uint32_t *__restrict__ len,
                                              const uint32_t *__restrict__ wl, int z, int own) {
    uint32_t ch = 0xFFFFFFFFu, t = WINF;
    for (int i0 = 0; i0 < n; i0 += WIN_BATCH) {
        uint32_t v[WIN_BATCH], w[WIN_BATCH];
#pragma unroll
        for (int k = 0; k < WIN_BATCH; k++) {
            const int j = i0 + k;
            const int i = DIR > 0 ? j : n - 1 - j;
            v[k] = j < n ? p[i] : 0xFFFFFFFFu;
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
                if (nd < dcur) {
                    const uint32_t nv = (v[k] & ~WMASK) | nd;
                    p[i] = nv; ch = xr_min(ch, nd);
                    if (AXIS == 0) { c.posd[own * c.WX + i] = 1; win2_push_from_A<C>(cluster, c, z, own, i, nv); }
                    else win2_push_from_B<C>(cluster, c, z, own, i, nv);
                }
                t = nd;
            }
        }
    }
    return ch;
}

// Warp-cooperative exact relaxation of one whole line (both directions) as two min-plus scans
// (see win_scan_line of the band kernel); lowered cells are pushed to the other layout.
template <int C, int AXIS>
__device__ __forceinline__ uint32_t win2_scan_line(cg::cluster_group &cluster, const Win2Ctx &c, uint32_t *__restrict__ p, int n,
                                                   int q, int g, uint32_t lutreg, uint32_t pen,
                                                   const uint32_t *__restrict__ len, const uint32_t *__restrict__ wl,
                                                   bool uni, int z, int own) {
    constexpr int G = 32;
    uint32_t dv[WIN_QMAX], wf[WIN_QMAX], wb[WIN_QMAX], fl[WIN_QMAX];
    const int i0 = g * q;
    unsigned chg = 0, valid = 0;
#pragma unroll
    for (int k = 0; k < WIN_QMAX; k++) {
        const int i = i0 + k;
        const bool ok = k < q && i < n;
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
    chg &= valid;
    uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < WIN_QMAX; k++) {
        if ((chg >> k) & 1u) {
            const int i = i0 + k;
            const uint32_t nv = fl[k] | dv[k];
            p[i] = nv;
            if (AXIS == 0) { c.posd[own * c.WX + i] = 1; win2_push_from_A<C>(cluster, c, z, own, i, nv); }
            else win2_push_from_B<C>(cluster, c, z, own, i, nv);
            mn = xr_min(mn, dv[k]);
        }
    }
    return mn;
}

#ifndef WIN2_COOP_LEN
#define WIN2_COOP_LEN 48             // lines of at least this many cells (and at most 32*WIN_QMAX) are scanned by a whole warp
#endif
template <int C>
__device__ uint32_t win2_sweep_y(cg::cluster_group &cluster, const Win2Ctx &c, int n, long long &work) {
    uint32_t ch = 0xFFFFFFFFu;
    if (c.WY >= WIN2_COOP_LEN && c.WY <= 32 * WIN_QMAX) {
        const int q = ((c.WY + 31) / 32) | 1;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int k = warp; k < n; k += WIN_T / 32) {
            const int col = c.list[k];
            const int z = col / c.WB, xb = col - z * c.WB;
            uint32_t *p = c.B + ((size_t)z * c.WB + xb) * c.WYp;
            ch = xr_min(ch, win2_scan_line<C, 1>(cluster, c, p, c.WY, q, lane, c.lutm[1 * c.Z + z], c.pens[z], c.leny,
                                                 c.wlut + (1 * c.Z + z) * 8, c.uni_y != 0, z, xb));
            if (lane == 0) work += c.WY;
        }
        return ch;
    }
    for (int k = threadIdx.x; k < n; k += WIN_T) {
        const int col = c.list[k];
        const int z = col / c.WB, xb = col - z * c.WB;
        uint32_t *p = c.B + ((size_t)z * c.WB + xb) * c.WYp;
        const uint32_t lutreg = c.lutm[1 * c.Z + z], pen = c.pens[z];
        const uint32_t *wl = c.wlut + (1 * c.Z + z) * 8;
        if (c.uni_y) {
            ch = xr_min(ch, win2_walk<C, 1, 1, true>(cluster, c, p, c.WY, lutreg, pen, c.leny, wl, z, xb));
            ch = xr_min(ch, win2_walk<C, 1, -1, true>(cluster, c, p, c.WY, lutreg, pen, c.leny, wl, z, xb));
        } else {
            ch = xr_min(ch, win2_walk<C, 1, 1, false>(cluster, c, p, c.WY, lutreg, pen, c.leny, wl, z, xb));
            ch = xr_min(ch, win2_walk<C, 1, -1, false>(cluster, c, p, c.WY, lutreg, pen, c.leny, wl, z, xb));
        }
        work += c.WY;
    }
    return ch;
}

template <int C>
__device__ uint32_t win2_sweep_x(cg::cluster_group &cluster, const Win2Ctx &c, int n, long long &work) {
    uint32_t ch = 0xFFFFFFFFu;
    if (c.WX >= WIN2_COOP_LEN && c.WX <= 32 * WIN_QMAX) {
        const int q = ((c.WX + 31) / 32) | 1;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int k = warp; k < n; k += WIN_T / 32) {
            const int row = c.list[k];
            const int z = row / c.HA, ya = row - z * c.HA;
            uint32_t *p = c.A + ((size_t)z * c.HA + ya) * c.WXp;
            ch = xr_min(ch, win2_scan_line<C, 0>(cluster, c, p, c.WX, q, lane, c.lutm[0 * c.Z + z], c.pens[z], c.lenx,
                                                 c.wlut + (0 * c.Z + z) * 8, c.uni_x != 0, z, ya));
            if (lane == 0) work += c.WX;
        }
        return ch;
    }
    for (int k = threadIdx.x; k < n; k += WIN_T) {
        const int row = c.list[k];
        const int z = row / c.HA, ya = row - z * c.HA;
        uint32_t *p = c.A + ((size_t)z * c.HA + ya) * c.WXp;
        const uint32_t lutreg = c.lutm[0 * c.Z + z], pen = c.pens[z];
        const uint32_t *wl = c.wlut + (0 * c.Z + z) * 8;
        if (c.uni_x) {
            ch = xr_min(ch, win2_walk<C, 0, 1, true>(cluster, c, p, c.WX, lutreg, pen, c.lenx, wl, z, ya));
            ch = xr_min(ch, win2_walk<C, 0, -1, true>(cluster, c, p, c.WX, lutreg, pen, c.lenx, wl, z, ya));
        } else {
            ch = xr_min(ch, win2_walk<C, 0, 1, false>(cluster, c, p, c.WX, lutreg, pen, c.lenx, wl, z, ya));
            ch = xr_min(ch, win2_walk<C, 0, -1, false>(cluster, c, p, c.WX, lutreg, pen, c.lenx, wl, z, ya));
        }
        work += c.WX;
    }
    return ch;
}

// via stacks of layout A (one thread per dirty (ya, x) position)
template <int C>
__device__ uint32_t win2_sweep_z(cg::cluster_group &cluster, const Win2Ctx &c, long long &work) {
    uint32_t ch = 0xFFFFFFFFu;
    const int npos = c.WX * c.ha;
    const size_t zs = (size_t)c.HA * c.WXp;
    const uint32_t lutreg = c.lutm[2 * c.Z];
    for (int pos = threadIdx.x; pos < npos; pos += WIN_T) {
        if (!c.posd[pos]) continue;
        c.posd[pos] = 0;
        const int ya = pos / c.WX, x = pos - ya * c.WX;
        uint32_t *p = c.A + (size_t)ya * c.WXp + x;
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
                c.rowd[z * c.HA + ya] = 1;
                win2_push_from_A<C>(cluster, c, z, ya, x, v[z]);
            }
        }
        work += c.Z;
    }
    return ch;
}

// pointer to the layout-A copy of window cell (lx, wy, z) in the CTA that owns it
template <int C>
__device__ __forceinline__ uint32_t *win2_cellA(cg::cluster_group &cluster, const Win2Ctx &c, int lx, int wy, int z) {
    uint32_t *p = c.A + ((size_t)z * c.HA + (wy >> Log2C<C>::v)) * c.WXp + lx;
    return C == 1 ? p : cluster.map_shared_rank(p, wy & (C - 1));
}
// Put window cell (lx, wy, z) on the tree in both layouts and flag its three lines.
template <int C>
__device__ __forceinline__ void win2_set_tree(cg::cluster_group &cluster, const Win2Ctx &c, int lx, int wy, int z, uint32_t flags) {
    const uint32_t nv = (flags & ~WMASK) | (CF_TREE << 28);
    const int ra = wy & (C - 1), ya = wy >> Log2C<C>::v, rb = lx & (C - 1), xb = lx >> Log2C<C>::v;
    uint32_t *pa = c.A + ((size_t)z * c.HA + ya) * c.WXp + lx, *pb = c.B + ((size_t)z * c.WB + xb) * c.WYp + wy;
    uint8_t *pr = c.rowd + z * c.HA + ya, *pp = c.posd + ya * c.WX + lx, *pc = c.cold + z * c.WB + xb;
    if (C > 1) {
        pa = cluster.map_shared_rank(pa, ra); pr = cluster.map_shared_rank(pr, ra); pp = cluster.map_shared_rank(pp, ra);
        pb = cluster.map_shared_rank(pb, rb); pc = cluster.map_shared_rank(pc, rb);
    }
    *pa = nv; *pb = nv; *pr = 1; *pp = 1; *pc = 1;
}

template <int C>
__global__ void __launch_bounds__(WIN_T, 1) k_route_win2(Geo g, Dev d, const int *env_list) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int env = env_list[blockIdx.x / C];
    const int net = d.act[2 * env + 1];
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int LC = Log2C<C>::v;
    const int *wd = d.net_win + ((size_t)env * (g.max_nets + 1) + net) * 6;
    const int wx0 = wd[0] & 0xFFFF, wy0 = wd[0] >> 16, WX = wd[1] & 0xFFFF, WY = wd[1] >> 16;
    Win2Ctx c;
    c.Z = g.Z; c.WX = WX; c.WXp = WX | 1; c.WY = WY; c.WYp = WY | 1; c.rank = rank;
    c.HA = (WY + C - 1) / C; c.WB = (WX + C - 1) / C;
    c.ha = WY > rank ? (WY - rank + C - 1) / C : 0;
    c.wb = WX > rank ? (WX - rank + C - 1) / C : 0;
    extern __shared__ __align__(16) uint32_t wsm[];
    c.A = wsm;
    c.B = c.A + (size_t)c.Z * c.HA * c.WXp;
    uint32_t *aux = c.B + (size_t)c.Z * c.WB * c.WYp;
    c.lutm = aux; aux += 3 * c.Z;
    c.pens = aux; aux += 3 * c.Z;
    c.wlut = aux; aux += 24 * c.Z;
    c.uni_x = g.uniform_x; c.uni_y = g.uniform_y;
    c.lenx = aux; aux += WX + 2;
    c.leny = aux; aux += WY + 2;
    aux += (aux - wsm) & 1;
    unsigned long long *s_best = reinterpret_cast<unsigned long long *>(aux); aux += 4;
    int *s_flag = reinterpret_cast<int *>(aux); aux += 8;   // [2] exit, [3] more, [4] #targets, [5] list count, [6] #source APs
    int *s_tgt = reinterpret_cast<int *>(aux); aux += 2 * WIN_TGT_CAP;
    uint32_t *s_tloc = aux; aux += WIN_TGT_CAP;             // layout-A index of the unconnected APs owned by this CTA
    uint32_t *s_red = aux; aux += 8;
    uint32_t *s_apc = aux; aux += WIN_TGT_CAP;              // access point i of the net: lx | wy << 10 | z << 20
    uint16_t *s_appin = reinterpret_cast<uint16_t *>(aux); aux += WIN_TGT_CAP / 2;
    uint16_t *s_tidx = reinterpret_cast<uint16_t *>(aux); aux += WIN_TGT_CAP / 2;    // unconnected access points
    uint8_t *s_apconn = reinterpret_cast<uint8_t *>(aux); aux += WIN_TGT_CAP / 4;
    uint8_t *s_apon = reinterpret_cast<uint8_t *>(aux); aux += WIN_TGT_CAP / 4;
    c.cnt = &s_flag[5];
    c.rowd = reinterpret_cast<uint8_t *>(aux);
    c.cold = c.rowd + c.Z * c.HA;
    c.posd = c.cold + c.Z * c.WB;
    const int n_flag_bytes = c.Z * c.HA + c.Z * c.WB + c.HA * WX;
    c.list = reinterpret_cast<uint16_t *>(c.rowd + (((size_t)n_flag_bytes + 3) & ~(size_t)3));
    c.sA = (uint32_t)__cvta_generic_to_shared(c.A); c.sB = (uint32_t)__cvta_generic_to_shared(c.B);
    c.sRowd = (uint32_t)__cvta_generic_to_shared(c.rowd); c.sCold = (uint32_t)__cvta_generic_to_shared(c.cold);
    c.sPosd = (uint32_t)__cvta_generic_to_shared(c.posd);
    // ---- tables (as in the band kernel; leny covers the whole window height)
    for (int i = tid; i < 3 * c.Z; i += WIN_T) {
        const int axis = i / c.Z, z = i - axis * c.Z;
        uint32_t r = 0;
        for (int f = 0; f < 4; f++)
            r |= (axis == 0 ? g.multX[z][f] : axis == 1 ? g.multY[z][f] : g.multV[f]) << (8 * f);
        c.lutm[i] = r;
        c.pens[i] = axis == 0 ? g.pen[z] : axis == 1 ? (z >= 1 ? g.vlen[z - 1] : 0u) : g.vlen[z];
    }
    for (int i = tid; i < 24 * c.Z; i += WIN_T) {
        const int axis = i / (8 * c.Z), z = (i / 8) % c.Z, f = i & 7;
        const uint32_t len = axis == 0 ? (uint32_t)g.dx : axis == 1 ? (uint32_t)g.dy : 0u;
        const uint32_t mult = axis == 0 ? g.multX[z][f & 3] : axis == 1 ? g.multY[z][f & 3] : g.multV[f & 3];
        c.wlut[i] = len * mult + ((f & 4) ? g.pen[z] : 0u);
    }
    for (int i = tid; i <= WX; i += WIN_T) {
        const int gx = wx0 + i;
        c.lenx[i] = (gx >= 1 && gx < g.X) ? (uint32_t)(g.xc[gx] - g.xc[gx - 1]) : 0u;
    }
    for (int i = tid; i <= WY; i += WIN_T) {
        const int gy = wy0 + i;
        c.leny[i] = (gy >= 1 && gy < g.Y) ? (uint32_t)(g.yc[gy] - g.yc[gy - 1]) : 0u;
    }
    for (int i = tid; i < n_flag_bytes; i += WIN_T) c.rowd[i] = 0;
    // ---- load both layouts: cost flags frozen now from the occupancy and the access-point owners, dist = INF
    const size_t eoff = (size_t)env * g.cells_p;
    for (int i = tid; i < c.Z * c.HA * c.WXp; i += WIN_T) {
        const int x = i % c.WXp, ya = (i / c.WXp) % c.HA, z = i / (c.WXp * c.HA);
        uint32_t v = WINF;
        if (x < WX && ya < c.ha) {
            const size_t gi = eoff + ((size_t)z * g.Y + wy0 + rank + C * ya) * g.Xp + wx0 + x;
            v |= cost_flags(__ldg(d.cellinfo + gi), __ldg(d.apnet + gi), (uint32_t)net) << 28;
        }
        c.A[i] = v;
    }
    for (int i = tid; i < c.Z * c.WB * c.WYp; i += WIN_T) {
        const int y = i % c.WYp, xb = (i / c.WYp) % c.WB, z = i / (c.WYp * c.WB);
        uint32_t v = WINF;
        if (y < WY && xb < c.wb) {
            const size_t gi = eoff + ((size_t)z * g.Y + wy0 + y) * g.Xp + wx0 + rank + C * xb;
            v |= cost_flags(__ldg(d.cellinfo + gi), __ldg(d.apnet + gi), (uint32_t)net) << 28;
        }
        c.B[i] = v;
    }
    if (tid < 8) s_flag[tid] = 0;
    __syncthreads();
    // ---- the net's access points, cached on chip (the host sends a net here only if it has at most
    // WIN_TGT_CAP of them): window cell, DBU position (exit-test heuristic), pin, connected flag
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const int s = ns[net], t = ns[net + 1], n_ap = t - s;
    const size_t aoff = (size_t)env * g.max_aps + s;
    const unsigned srcpin = d.net_srcpin[(size_t)env * (g.max_nets + 1) + net];
    for (int i = tid; i < n_ap; i += WIN_T) {
        const int cp = d.ap_cellp[aoff + i];
        const int gx = cp % g.Xp, gy = (cp / g.Xp) % g.Y, z = cp / (g.Xp * g.Y);
        const unsigned pin = d.ap_pin[aoff + i];
        s_apc[i] = (uint32_t)((gx - wx0) | ((gy - wy0) << 10) | (z << 20));
        s_tgt[2 * i] = g.xc[gx]; s_tgt[2 * i + 1] = g.yc[gy];
        s_appin[i] = (uint16_t)pin;
        s_apconn[i] = pin == srcpin;
        if (pin != srcpin) continue;
        atomicAdd(&s_flag[6], 1);
        const int x = gx - wx0, wy = gy - wy0;               // seeds, in the layouts that own them
        if ((wy & (C - 1)) == rank) {
            const int ya = wy >> LC;
            c.A[((size_t)z * c.HA + ya) * c.WXp + x] &= ~WMASK;
            c.rowd[z * c.HA + ya] = 1; c.posd[ya * WX + x] = 1;
        }
        if ((x & (C - 1)) == rank) {
            const int xb = x >> LC;
            c.B[((size_t)z * c.WB + xb) * c.WYp + wy] &= ~WMASK;
            c.cold[z * c.WB + xb] = 1;
        }
    }
    bool first = true;
    long long work = 0, cyc_relax = 0;
    int n_iter = 0, n_conn = 0;
#ifdef WIN_PHASE_TIMING
    long long ph[7] = {0, 0, 0, 0, 0, 0, 0};   // y, wait after y, x, z, closing wait, dirty columns, dirty rows
#endif
    const long long tk0 = clock64();
    const bool open_x0 = wx0 > 0, open_x1 = wx0 + WX < g.X, open_y0 = wy0 > 0, open_y1 = wy0 + WY < g.Y;
    int parity = 0;
    if (C > 1) cluster.sync(); else __syncthreads();
    const int n_src_ap = s_flag[6];
    for (;;) {                                            // ---- one connection per trip
        // ---- targets: the unconnected access points (s_tidx), those this CTA owns in layout A (s_tloc)
        if (tid == 0) { s_flag[4] = 0; s_red[4] = 0; }
        __syncthreads();
        for (int i = tid; i < n_ap; i += WIN_T) {
            if (s_apconn[i]) continue;
            const uint32_t pc = s_apc[i];
            const int x = pc & 1023, wy = (pc >> 10) & 1023, z = pc >> 20;
            s_tidx[atomicAdd(&s_flag[4], 1)] = (uint16_t)i;
            if ((wy & (C - 1)) == rank)
                s_tloc[atomicAdd(&s_red[4], 1u)] = (uint32_t)(((size_t)z * c.HA + (wy >> LC)) * c.WXp + x);
        }
        __syncthreads();
        const int n_tgt = s_flag[4];
        const int n_loc = (int)s_red[4];
        const long long tr0 = clock64();
        for (;;) {                                        // ---- relax (bounded early stop as in the band kernel)
            n_iter++;
            if (tid == 0) { s_red[2 * parity] = 0xFFFFFFFFu; s_red[2 * parity + 1] = 0xFFFFFFFFu; }
#ifdef WIN_PHASE_TIMING
            const long long q0 = clock64();
#endif
            const int ny = win2_compact(c, c.cold, c.Z * c.WB);
            uint32_t ch = win2_sweep_y<C>(cluster, c, ny, work);
#ifdef WIN_PHASE_TIMING
            __syncthreads();
            const long long q1 = clock64();
#endif
            if (C > 1) cluster.sync(); else __syncthreads();      // pushes of the y phase have landed in layout A
#ifdef WIN_PHASE_TIMING
            const long long q2 = clock64();
#endif
            const int nx = win2_compact(c, c.rowd, c.Z * c.HA);
            ch = xr_min(ch, win2_sweep_x<C>(cluster, c, nx, work));
            __syncthreads();
#ifdef WIN_PHASE_TIMING
            const long long q3 = clock64();
#endif
            ch = xr_min(ch, win2_sweep_z<C>(cluster, c, work));
#ifdef WIN_PHASE_TIMING
            __syncthreads();
            const long long q4 = clock64();
            ph[0] += q1 - q0; ph[1] += q2 - q1; ph[2] += q3 - q2; ph[3] += q4 - q3; ph[5] += ny; ph[6] += nx;
#endif
            ch = __reduce_min_sync(0xFFFFFFFFu, ch);
            if (lane == 0 && ch != 0xFFFFFFFFu) atomicMin(&s_red[2 * parity], ch);
            __syncthreads();
            {
                uint32_t bl = 0xFFFFFFFFu;
                for (int k = tid; k < n_loc; k += WIN_T) bl = xr_min(bl, c.A[s_tloc[k]] & WMASK);
                bl = __reduce_min_sync(0xFFFFFFFFu, bl);
                if (lane == 0 && bl != 0xFFFFFFFFu) atomicMin(&s_red[2 * parity + 1], bl);
            }
#ifdef WIN_PHASE_TIMING
            const long long q5 = clock64();
#endif
            if (C > 1) cluster.sync(); else __syncthreads();      // pushes of the x and via phases have landed in layout B
#ifdef WIN_PHASE_TIMING
            ph[4] += clock64() - q5;
#endif
            uint32_t gmin = 0xFFFFFFFFu, gB = 0xFFFFFFFFu;
            for (int r = 0; r < C; r++) {
                const uint32_t *rr = (C > 1) ? cluster.map_shared_rank(&s_red[2 * parity], r) : &s_red[2 * parity];
                gmin = xr_min(gmin, rr[0]); gB = xr_min(gB, rr[1]);
            }
            parity ^= 1;
            if (gmin == 0xFFFFFFFFu) break;
            if (gB < WINF && gmin >= gB) break;
        }
        const long long tq0 = clock64();
        cyc_relax += tq0 - tr0; n_conn++;
        // ---- best target among the access points this CTA owns in layout A: argmin (dist, padded cell index)
        if (tid == 0) s_best[0] = ~0ull;
        __syncthreads();
        unsigned long long best = ~0ull;
        for (int k = tid; k < n_tgt; k += WIN_T) {
            const uint32_t pc = s_apc[s_tidx[k]];
            const int x = pc & 1023, wy = (pc >> 10) & 1023, z = pc >> 20;
            if ((wy & (C - 1)) != rank) continue;
            const uint32_t dv = c.A[((size_t)z * c.HA + (wy >> LC)) * c.WXp + x] & WMASK;
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
        // ---- window-exit test over the open faces (cells of layout A owned by this CTA)
        bool esc = (best == ~0ull) || B >= WINF;
        if (!esc && (open_x0 || open_x1 || open_y0 || open_y1)) {
            const int nface = c.Z * c.ha * 2 + c.Z * WX * 2;
            for (int i = tid; i < nface && !esc; i += WIN_T) {
                int x, ya, z; bool open;
                if (i < c.Z * c.ha * 2) {
                    const int side = i & 1, k = i >> 1; z = k / c.ha; ya = k - z * c.ha;
                    x = side ? WX - 1 : 0; open = side ? open_x1 : open_x0;
                } else {
                    const int k0 = i - c.Z * c.ha * 2; const int side = k0 & 1, k = k0 >> 1; z = k / WX; x = k - z * WX;
                    const int wy = side ? WY - 1 : 0;
                    ya = wy >> LC;
                    open = (side ? open_y1 : open_y0) && (wy & (C - 1)) == rank;
                }
                if (!open) continue;
                const uint32_t dv = c.A[((size_t)z * c.HA + ya) * c.WXp + x] & WMASK;
                if (dv >= WINF || dv > B) continue;
                const int px = g.xc[wx0 + x], py = g.yc[wy0 + rank + C * ya];
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
        // ---- canonical backtrace by warp 0 of rank 0 (layout A read over DSMEM).  The walk only touches the
        // on-chip copies and the path list; the global state (occupancy, observation bytes, tree flags) of all
        // new path cells is committed afterwards in one parallel pass.
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
            int pend = 0;                                // cells waiting for their global commit (s_tloc is free here)
            auto flush = [&]() {
                __syncwarp();
                for (int k = lane; k < pend; k += 32) {
                    const uint32_t pc = s_tloc[k];
                    commit_cell(g, d, env, net, wx0 + (int)(pc & 1023u), wy0 + (int)((pc >> 10) & 1023u), (int)(pc >> 20));
                }
                __syncwarp();
                pend = 0;
            };
            for (;;) {
                __syncwarp();
                const uint32_t vc = *win2_cellA<C>(cluster, c, cx - wx0, cy - wy0, cz);
                const uint32_t dc = vc & WMASK;
                if (dc == 0) break;
                if (last >= 0) {
                    int ddx, ddy, ddz; dir_delta(last, ddx, ddy, ddz);
                    const int ax = cx - lane * ddx, ay = cy - lane * ddy, az = cz - lane * ddz;
                    const int bx = ax - ddx, by = ay - ddy, bz = az - ddz;
                    bool ok = inwin(ax, ay, az) && inwin(bx, by, bz);
                    uint32_t va = 0;
                    if (ok) {
                        va = *win2_cellA<C>(cluster, c, ax - wx0, ay - wy0, az);
                        const uint32_t vb = *win2_cellA<C>(cluster, c, bx - wx0, by - wy0, bz);
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
                        const uint32_t dp = *win2_cellA<C>(cluster, c, px - wx0, py - wy0, pz) & WMASK;
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
            if (!fail) {                                 // the cell the walk ended on joins the tree on the first connection only
                if (first && pend + 1 > WIN_TGT_CAP) flush();
                if (lane == 0) {
                    if (first) s_tloc[pend] = (uint32_t)((cx - wx0) | ((cy - wy0) << 10) | (cz << 20));
                    if (pn < g.path_cap) path[pn] = (cz * g.Y + cy) * g.X + cx;
                }
                pn += 1;
                if (first) pend += 1;
            }
            flush();
            // The walk is over: only now do its cells join the tree on chip, in both layouts (distance 0, tree bit, dirty
            // lines) -- a cell zeroed under the walk could pass the predecessor test of a later cell.  They are read
            // back from the path record, which must therefore hold the whole net (an overflow fails the step with
            // XR_E_CAPACITY).  The cell the walk ended on is a source already; on the first connection it gets its
            // tree bit here.
            const bool over = pn > g.path_cap;
            for (int k = pn0 + lane; k < (over ? g.path_cap : pn); k += 32) {
                const int ci = path[k];
                const int x = ci % g.X - wx0, y = (ci / g.X) % g.Y - wy0, z = ci / (g.X * g.Y);
                win2_set_tree<C>(cluster, c, x, y, z, *win2_cellA<C>(cluster, c, x, y, z));
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
        // ---- pin bookkeeping, by every CTA on its own copy: a pin is connected once any of its access points is
        // on the tree (tree bit of the on-chip cell); rank 0 mirrors the flags to global memory for a hand-over
        for (int i = tid; i < n_ap; i += WIN_T) {
            const uint32_t pc = s_apc[i];
            s_apon[i] = s_apconn[i] ? 1 : (((*win2_cellA<C>(cluster, c, pc & 1023, (pc >> 10) & 1023, pc >> 20)) >> 31) & 1u);
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
        if (!more) break;
        if (first) {
            // after the first connection only the path is the tree (see the band kernel)
            if (n_src_ap > 1) {
                for (int i = tid; i < c.Z * c.HA * c.WXp; i += WIN_T) {
                    const uint32_t v = c.A[i];
                    const bool tree = ((v >> 28) & CF_TREE) != 0;
                    c.A[i] = (v & ~WMASK) | (tree ? 0u : WINF);
                    if (tree) {
                        const int x = i % c.WXp, ya = (i / c.WXp) % c.HA, z = i / (c.WXp * c.HA);
                        if (x < WX && ya < c.ha) { c.rowd[z * c.HA + ya] = 1; c.posd[ya * WX + x] = 1; }
                    }
                }
                for (int i = tid; i < c.Z * c.WB * c.WYp; i += WIN_T) {
                    const uint32_t v = c.B[i];
                    const bool tree = ((v >> 28) & CF_TREE) != 0;
                    c.B[i] = (v & ~WMASK) | (tree ? 0u : WINF);
                    if (tree) {
                        const int y = i % c.WYp, xb = (i / c.WYp) % c.WB, z = i / (c.WYp * c.WB);
                        if (y < WY && xb < c.wb) c.cold[z * c.WB + xb] = 1;
                    }
                }
                if (C > 1) cluster.sync(); else __syncthreads();  // nobody pushes into a copy that is still being rebuilt
            }
            first = false;
        }
    }
    if (tid == 0 && rank == 0 && d.dbg) {
        atomicAdd(&d.dbg[0], (unsigned long long)n_iter); atomicAdd(&d.dbg[1], (unsigned long long)n_conn);
        atomicAdd(&d.dbg[2], (unsigned long long)cyc_relax); atomicAdd(&d.dbg[3], (unsigned long long)(clock64() - tk0));
        atomicAdd(&d.dbg[4], 1ull); atomicAdd(&d.dbg[5], (unsigned long long)(WX * WY));
        atomicAdd(&d.dbg[6], (unsigned long long)n_iter); atomicAdd(&d.dbg[7], (unsigned long long)cyc_relax);
        atomicAdd(&d.dbg[15], (unsigned long long)n_conn);
#ifdef WIN_PHASE_TIMING
        for (int k = 0; k < 7; k++) atomicAdd(&d.dbg[8 + k], (unsigned long long)ph[k]);
#endif
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) work += __shfl_xor_sync(0xFFFFFFFFu, work, off);
    if (lane == 0 && work)
        atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 7]), (unsigned long long)work);
    if (tid == 0 && rank == 0)
        atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 2]), (unsigned long long)(3 * n_iter));
    if (C > 1) cluster.sync();
}
