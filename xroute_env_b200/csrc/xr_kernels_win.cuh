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
#define WIN_TGT_CAP 256
#define WIN_AUX_WORDS(Z, HH, WX) (3 * (Z) * 8 + (WX) + 2 + (HH) + 2 + 64 + 2 * WIN_TGT_CAP)

struct WinCtx {
    uint32_t *cell;      // [Z][HH][WXp]
    uint32_t *lut;       // [3][Z][8]  mult | pen << 8
    uint32_t *lenx;      // [WX+1]  lenx[lx] = xc[wx0+lx] - xc[wx0+lx-1]
    uint32_t *leny;      // [HH+1]  leny[ly] = yc[gy] - yc[gy-1], gy = wy0 + ry0 + ly - 1
    int Z, WX, WXp, HH, h; // h = real rows of this CTA (local rows 1..h)
    int wx0, wy0, ry0;
};

__device__ __forceinline__ uint32_t win_w(const uint32_t *lutrow, uint32_t len, uint32_t f) {
    const uint32_t e = lutrow[f & 7u];
    return len * (e & 0xFFu) + (e >> 8);
}

// one thread per (x, z) column of the band: forward from the upper halo, back from the lower
__device__ bool win_sweep_y(const WinCtx &c) {
    bool ch = false;
    const int ncol = c.WX * c.Z;
    for (int col = threadIdx.x; col < ncol; col += WIN_T) {
        const int z = col / c.WX, x = col - z * c.WX;
        uint32_t *p = c.cell + (size_t)z * c.HH * c.WXp + x;
        const uint32_t *lr = c.lut + (1 * c.Z + z) * 8;
        uint32_t t = p[0] & WMASK;
        for (int ly = 1; ly <= c.h; ly++) {
            const uint32_t v = p[ly * c.WXp];
            const uint32_t dcur = v & WMASK;
            const uint32_t nd = xr_min(t + win_w(lr, c.leny[ly], v >> 28), dcur);
            if (nd < dcur) { p[ly * c.WXp] = (v & ~WMASK) | nd; ch = true; }
            t = nd;
        }
        t = p[(c.h + 1) * c.WXp] & WMASK;
        for (int ly = c.h; ly >= 1; ly--) {
            const uint32_t v = p[ly * c.WXp];
            const uint32_t dcur = v & WMASK;
            const uint32_t nd = xr_min(t + win_w(lr, c.leny[ly + 1], v >> 28), dcur);
            if (nd < dcur) { p[ly * c.WXp] = (v & ~WMASK) | nd; ch = true; }
            t = nd;
        }
    }
    return ch;
}

// one thread per (ly, z) row of the band
__device__ bool win_sweep_x(const WinCtx &c) {
    bool ch = false;
    const int nrow = c.h * c.Z;
    for (int row = threadIdx.x; row < nrow; row += WIN_T) {
        const int z = row / c.h, ly = row - z * c.h + 1;
        uint32_t *p = c.cell + ((size_t)z * c.HH + ly) * c.WXp;
        const uint32_t *lr = c.lut + (0 * c.Z + z) * 8;
        uint32_t t = WINF;
        for (int x = 0; x < c.WX; x++) {
            const uint32_t v = p[x];
            const uint32_t dcur = v & WMASK;
            const uint32_t nd = xr_min(t + win_w(lr, c.lenx[x], v >> 28), dcur);
            if (nd < dcur) { p[x] = (v & ~WMASK) | nd; ch = true; }
            t = nd;
        }
        t = WINF;
        for (int x = c.WX - 1; x >= 0; x--) {
            const uint32_t v = p[x];
            const uint32_t dcur = v & WMASK;
            const uint32_t nd = xr_min(t + win_w(lr, c.lenx[x + 1], v >> 28), dcur);
            if (nd < dcur) { p[x] = (v & ~WMASK) | nd; ch = true; }
            t = nd;
        }
    }
    return ch;
}

// one thread per (x, ly) position: up then down through the layers
__device__ bool win_sweep_z(const WinCtx &c, const Geo &g) {
    bool ch = false;
    const int npos = c.WX * c.h;
    const size_t zs = (size_t)c.HH * c.WXp;
    for (int pos = threadIdx.x; pos < npos; pos += WIN_T) {
        const int lyi = pos / c.WX, x = pos - lyi * c.WX;
        uint32_t *p = c.cell + (size_t)(lyi + 1) * c.WXp + x;
        uint32_t t = p[0] & WMASK;
        for (int z = 1; z < c.Z; z++) {
            const uint32_t v = p[z * zs];
            const uint32_t dcur = v & WMASK;
            const uint32_t nd = xr_min(t + win_w(c.lut + (2 * c.Z + z) * 8, g.vlen[z - 1], v >> 28), dcur);
            if (nd < dcur) { p[z * zs] = (v & ~WMASK) | nd; ch = true; }
            t = nd;
        }
        for (int z = c.Z - 2; z >= 0; z--) {
            const uint32_t v = p[z * zs];
            const uint32_t dcur = v & WMASK;
            const uint32_t nd = xr_min(t + win_w(c.lut + (2 * c.Z + z) * 8, g.vlen[z], v >> 28), dcur);
            if (nd < dcur) { p[z * zs] = (v & ~WMASK) | nd; ch = true; }
            t = nd;
        }
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

template <int C>
__global__ void __launch_bounds__(WIN_T, 1) k_route_win(Geo g, Dev d, const int *env_list) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (C > 1) ? (int)cluster.block_rank() : 0;
    const int env = env_list[blockIdx.x / C];
    const int net = d.act[2 * env + 1];
    const int tid = threadIdx.x, lane = tid & 31;
    const int *wd = d.net_win + ((size_t)env * (g.max_nets + 1) + net) * 6;
    const int wx0 = wd[0] & 0xFFFF, wy0 = wd[0] >> 16, WX = wd[1] & 0xFFFF, WY = wd[1] >> 16;
    const int bx0 = wd[2], bx1 = wd[3], by0 = wd[4], by1 = wd[5];      // DBU bbox of the net's APs
    const int H = (WY + C - 1) / C;
    WinCtx c;
    c.Z = g.Z; c.WX = WX; c.WXp = WX | 1; c.HH = H + 2;
    c.wx0 = wx0; c.wy0 = wy0; c.ry0 = rank * H;
    c.h = WY - c.ry0; if (c.h > H) c.h = H; if (c.h < 0) c.h = 0;
    extern __shared__ __align__(16) uint32_t wsm[];
    c.cell = wsm;
    uint32_t *aux = wsm + (size_t)c.Z * c.HH * c.WXp;
    c.lut = aux; aux += 3 * c.Z * 8;
    c.lenx = aux; aux += WX + 2;
    c.leny = aux; aux += c.HH + 2;
    unsigned long long *s_best = reinterpret_cast<unsigned long long *>(aux + (aux - wsm) % 2);   // 8-byte aligned
    int *s_flag = reinterpret_cast<int *>(s_best + 2);   // [0..1] changed (double buffered), [2] exit-check, [3] state, [4] #targets
    int *s_tgt = s_flag + 8;                             // [WIN_TGT_CAP][2] DBU coordinates of the unconnected APs
    // ---- tables
    for (int i = tid; i < 3 * c.Z * 8; i += WIN_T) {
        const int axis = i / (c.Z * 8), z = (i / 8) % c.Z, f = i & 7;
        const uint32_t mult = axis == 0 ? g.multX[z][f & 3] : axis == 1 ? g.multY[z][f & 3] : g.multV[f & 3];
        c.lut[i] = mult | (((f & 4) ? g.pen[z] : 0u) << 8);
    }
    for (int i = tid; i <= WX; i += WIN_T) {
        const int gx = wx0 + i;
        c.lenx[i] = (gx >= 1 && gx < g.X) ? (uint32_t)(g.xc[gx] - g.xc[gx - 1]) : 0u;
    }
    for (int i = tid; i <= c.HH; i += WIN_T) {
        const int gy = wy0 + c.ry0 + i - 1;
        c.leny[i] = (gy >= 1 && gy < g.Y) ? (uint32_t)(g.yc[gy] - g.yc[gy - 1]) : 0u;
    }
    // ---- load the band: flags from the frozen cflag field, dist = INF; halo rows INF
    const size_t eoff = (size_t)env * g.cells_p;
    for (int i = tid; i < c.Z * c.HH * c.WXp; i += WIN_T) {
        const int x = i % c.WXp, ly = (i / c.WXp) % c.HH, z = i / (c.WXp * c.HH);
        uint32_t v = WINF;
        if (x < WX && ly >= 1 && ly <= c.h) {
            const int gy = wy0 + c.ry0 + ly - 1;
            v |= ((uint32_t)d.cflag[eoff + ((size_t)z * g.Y + gy) * g.Xp + wx0 + x] & 7u) << 28;
        }
        c.cell[i] = v;
    }
    if (tid < 8) s_flag[tid] = 0;
    __syncthreads();
    // ---- seeds: access points of the source pin inside this band
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const int s = ns[net], t = ns[net + 1];
    const size_t aoff = (size_t)env * g.max_aps;
    const unsigned srcpin = d.net_srcpin[(size_t)env * (g.max_nets + 1) + net];
    for (int i = s + tid; i < t; i += WIN_T) {
        if (d.ap_pin[aoff + i] != srcpin) continue;
        const int cp = d.ap_cellp[aoff + i];
        const int x = cp % g.Xp - wx0, wy = (cp / g.Xp) % g.Y - wy0, z = cp / (g.Xp * g.Y);
        const int ly = wy - c.ry0 + 1;
        if (ly >= 1 && ly <= c.h) c.cell[((size_t)z * c.HH + ly) * c.WXp + x] &= ~WMASK;
    }
    bool first = true;
    long long relaxed = 0, cyc_relax = 0;
    int n_iter = 0, n_conn = 0;
    const long long tk0 = clock64();
    const bool open_x0 = wx0 > 0, open_x1 = wx0 + WX < g.X, open_y0 = wy0 > 0, open_y1 = wy0 + WY < g.Y;
    const int band_cells = c.Z * c.h * WX;
    int parity = 0;
    if (C > 1) cluster.sync(); else __syncthreads();
    for (;;) {                                            // ---- one connection per trip
        // ---- relax to the fixpoint
        const long long tr0 = clock64();
        for (;;) {
            n_iter++;
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
                    c.cell[((size_t)z * c.HH + dst_ly) * c.WXp + x] = *rp;
                }
                cluster.sync();
            }
            bool ch = win_sweep_y(c);
            __syncthreads();
            ch |= win_sweep_x(c);
            __syncthreads();
            ch |= win_sweep_z(c, g);
            relaxed += 3ll * band_cells;
            const int anyc = __syncthreads_or(ch);
            if (C > 1) {
                if (tid == 0) s_flag[parity] = anyc;
                cluster.sync();
                int tot = 0;
                for (int r = 0; r < C; r++) tot |= *cluster.map_shared_rank(&s_flag[parity], r);
                parity ^= 1;
                if (!tot) break;
            } else if (!anyc) break;
        }
        // ---- best target in this band + window-exit test
        const long long tq0 = clock64();
        cyc_relax += tq0 - tr0; n_conn++;
        if (tid == 0) { s_best[0] = ~0ull; s_flag[4] = 0; }
        __syncthreads();
        unsigned long long best = ~0ull;
        for (int i = s + tid; i < t; i += WIN_T) {
            if (d.ap_conn[aoff + i]) continue;
            const int cp = d.ap_cellp[aoff + i];
            const int gx = cp % g.Xp, gy = (cp / g.Xp) % g.Y, z = cp / (g.Xp * g.Y);
            const int k = atomicAdd(&s_flag[4], 1);          // every CTA keeps the full target list
            if (k < WIN_TGT_CAP) { s_tgt[2 * k] = g.xc[gx]; s_tgt[2 * k + 1] = g.yc[gy]; }
            const int x = gx - wx0, wy = gy - wy0;
            const int ly = wy - c.ry0 + 1;
            if (ly < 1 || ly > c.h) continue;
            const uint32_t dv = c.cell[((size_t)z * c.HH + ly) * c.WXp + x] & WMASK;
            const unsigned long long key = ((unsigned long long)dv << 32) | (unsigned)cp;
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
        const int n_tgt = s_flag[4];
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
                uint32_t hmin;
                if (n_tgt <= WIN_TGT_CAP) {
                    hmin = 0xFFFFFFFFu;
                    for (int j = 0; j < n_tgt; j++) {
                        const uint32_t hh = (uint32_t)(abs(px - s_tgt[2 * j]) + abs(py - s_tgt[2 * j + 1]));
                        hmin = hh < hmin ? hh : hmin;
                    }
                } else {
                    const uint32_t hx = px < bx0 ? bx0 - px : (px > bx1 ? px - bx1 : 0);
                    const uint32_t hy = py < by0 ? by0 - py : (py > by1 ? py - by1 : 0);
                    hmin = hx + hy;
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
                    d.phase[env] = 1; d.changed[env] = 1; d.reinit[env] = first ? 0 : 1; d.first[env] = first ? 1 : 0;
                    atomicAdd(&d.flags[0], 1); atomicAdd(&d.flags[2], 1);
                }
                break;
            }
        } else if (esc_any) {
            if (tid == 0) {
                d.phase[env] = 1; d.changed[env] = 1; d.reinit[env] = first ? 0 : 1; d.first[env] = first ? 1 : 0;
                atomicAdd(&d.flags[0], 1); atomicAdd(&d.flags[2], 1);
            }
            break;
        }
        // ---- canonical backtrace + commit by warp 0 of rank 0 (cells read over DSMEM)
        if (rank == 0 && tid < 32) {
            int cp = (int)(best & 0xFFFFFFFFu);
            int cx = cp % g.Xp, cy = (cp / g.Xp) % g.Y, cz = cp / (g.Xp * g.Y);
            int pn = d.path_n[env];
            const int cn = d.conn_n[env];
            int *path = d.path + (size_t)env * g.path_cap;
            long long wl = 0, via = 0;
            int last = -1;
            bool fail = false;
            auto inwin = [&](int x, int y, int z) {
                return x >= wx0 && x < wx0 + WX && y >= wy0 && y < wy0 + WY && z >= 0 && z < g.Z;
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
                        if (lane < run) {
                            commit_cell(g, d, env, net, ax, ay, az);
                            *pa = (*pa & ~WMASK) | (CF_TREE << 28);
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
                if (lane == dir) {
                    commit_cell(g, d, env, net, cx, cy, cz);
                    *pc = (vc & ~WMASK) | (CF_TREE << 28);
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
            if (!fail) {
                if (lane == 0) {
                    if (first) {
                        commit_cell(g, d, env, net, cx, cy, cz);
                        uint32_t *pc = win_cell_ptr<C>(cluster, c, H, cx - wx0, cy - wy0, cz);
                        *pc = (*pc & ~WMASK) | (CF_TREE << 28);
                    }
                    if (pn < g.path_cap) path[pn] = (cz * g.Y + cy) * g.X + cx;
                }
                pn += 1;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                wl += __shfl_xor_sync(0xFFFFFFFFu, wl, off);
                via += __shfl_xor_sync(0xFFFFFFFFu, via, off);
            }
            __syncwarp();
            // pin bookkeeping on the global TREE marks written by commit_cell
            __threadfence_block();
            for (int i = s + lane; i < t; i += 32) {
                if (d.ap_conn[aoff + i]) continue;
                const unsigned pin = d.ap_pin[aoff + i];
                bool on = false;
                for (int j = i; j >= s && d.ap_pin[aoff + j] == pin && !on; j--)
                    on = (d.cflag[eoff + d.ap_cellp[aoff + j]] & CF_TREE) != 0;
                for (int j = i + 1; j < t && d.ap_pin[aoff + j] == pin && !on; j++)
                    on = (d.cflag[eoff + d.ap_cellp[aoff + j]] & CF_TREE) != 0;
                if (on) d.ap_conn[aoff + i] = 2;
            }
            __syncwarp();
            bool left = false;
            for (int i = s + lane; i < t; i += 32) {
                uint8_t v = d.ap_conn[aoff + i];
                if (v == 2) { d.ap_conn[aoff + i] = 1; v = 1; }
                left |= (v == 0);
            }
            left = __any_sync(0xFFFFFFFFu, left);
            if (lane == 0) {
                d.wlvia[2 * env] += wl; d.wlvia[2 * env + 1] += via;
                d.path_n[env] = pn;
                if (cn < g.conn_cap) {
                    d.conn_cost[(size_t)env * g.conn_cap + cn] = B;
                    d.conn_off[(size_t)env * (g.conn_cap + 1) + cn + 1] = pn;
                }
                d.conn_n[env] = cn + 1;
                d.envstat[8 * (size_t)env + 3] += 1;
                if (fail) d.flags[1] = 3;
                s_flag[3] = (left && !fail) ? 1 : 0;
                __threadfence();
            }
        }
        if (C > 1) cluster.sync(); else __syncthreads();
        const int more = (C > 1) ? *cluster.map_shared_rank(&s_flag[3], 0) : s_flag[3];
        if (!more) break;
        if (first) {
            // after the first connection only the path is the tree: the unused APs of the
            // source pin leave the source set
            for (int i = tid; i < c.Z * c.HH * c.WXp; i += WIN_T) {
                const uint32_t v = c.cell[i];
                c.cell[i] = (v & ~WMASK) | (((v >> 28) & CF_TREE) ? 0u : WINF);
            }
            first = false;
        }
        if (C > 1) cluster.sync(); else __syncthreads();
    }
    // relaxation accounting (cells touched by the in-window sweeps)
    if (tid == 0 && rank == 0 && d.dbg) {
        atomicAdd(&d.dbg[0], (unsigned long long)n_iter); atomicAdd(&d.dbg[1], (unsigned long long)n_conn);
        atomicAdd(&d.dbg[2], (unsigned long long)cyc_relax); atomicAdd(&d.dbg[3], (unsigned long long)(clock64() - tk0));
        atomicAdd(&d.dbg[4], 1ull); atomicAdd(&d.dbg[5], (unsigned long long)(WX * WY));
    }
    if (tid == 0) {                                        // every thread counted the same band
        atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 7]), (unsigned long long)relaxed);
        if (rank == 0 && band_cells > 0)
            atomicAdd(reinterpret_cast<unsigned long long *>(&d.envstat[8 * (size_t)env + 2]),
                      (unsigned long long)(relaxed / band_cells));
    }
    if (C > 1) cluster.sync();                             // keep peers' shared memory alive until all are done
}
