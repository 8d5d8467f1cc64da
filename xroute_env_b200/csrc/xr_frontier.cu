// xr_frontier.cu -- maze route of the selected net by a goal-directed frontier search (the default engine).
//
// Replaces the external TritonRoute maze search the reference reaches through ZMQ
// (baseline/baseline_utils.py:409-419); specification DESIGN.md section 3, same results as the sweep engines
// (xr_kernels_win*.cuh, xr_kernels_maze.cuh) and the oracle, bit for bit.
//
// One CTA routes one net, connection by connection.  A connection is a label-correcting best-first search from the
// whole tree on f = d + h, h = L1 track distance (DBU) to the nearest unconnected pin's access-point box: admissible and
// consistent because every planar move costs at least its length and vias cost > 0.  The search stops once no open
// entry has f <= B (best target distance found): every cell with d + h <= B then holds its final distance, which is
// all the target choice and the canonical walk read (tests/test_oracle_maze.py).  Measured on
// the bench workload that is 0.8-3 % of the cells a converged field would need.
//
//   field     dist64[cell] = (~epoch) << 32 | d in global memory (L2 resident): one 64-bit atomicMin relaxes a cell,
//             a new connection bumps the epoch, so the field is never cleared and has no window
//   open list (cell, f) entries, ping-pong, first cap_s entries in shared memory, the rest spilled to global memory
//   round     every entry with f <= fmin + delta is expanded at once (delta-stepping on f); an expansion relaxes a RAY
//             of up to `ray` cells along the layer's preferred direction, each way, plus the four other neighbours:
//             all loads of a ray are issued together, the prefix sums of the edge weights give the tentative
//             distances, the ray ends at the first cell it does not lower.  Rays cut the number of rounds (the serial
//             depth of the search) by 4-5x: a round costs one L2 round trip + two block barriers
//   epilogue  target = argmin (d, cell); canonical walk by warp 0 (32 cells of a straight run per round trip);
//             parallel commit; the new tree cells are marked in the field with the next epoch, which is also how the
//             pins that the path touched are found
#include "xr_frontier.h"

#include <climits>

#ifdef FR_TIMING
#define FR_TICK(k) do { const long long t__ = clock64(); ph[k] += t__ - tlast; tlast = t__; } while (0)
#else
#define FR_TICK(k) do { } while (0)
#endif

struct FrSm {
    int cnt[2];            // entries in the two open lists
    uint32_t fmin[2];      // smallest f pushed into each
    int nexp[2];           // single-step tasks of the round (double buffered counter)
    int nray[2];           // ray tasks of the round
    uint32_t B;            // best target distance seen so far in this connection
    int nbox;
    int err;
    int pn, cn, walk_fail;
    unsigned long long best;
    int rx0, rx1, ry0, ry1;  // tiles holding the net's access points
    int dmul;              // current bucket width in units of delta
    int far_n[2];          // entries parked in the two far lists (global memory only)
    uint32_t far_min[2];   // smallest f parked in each
};

__device__ __forceinline__ int fr_agg_inc(int *ctr) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

#define FR_RAY_POS 0x01000000u   // open-list entry: shoot a ray in the positive / negative preferred direction when expanded
#define FR_RAY_NEG 0x02000000u
#define FR_RAYS_BOTH 0x03000000u
#define FR_TILE_SHIFT 4    // heuristic tiles of 16 x 16 cells
#define FR_MAXTILES 4096

// WIDE: the variant for grids whose searches hold 10^5 open entries (far-list parking, task prefetch, batched
// classification loads); the plain variant has the shorter round for the small searches of small grids.
template <bool WIDE>
__global__ void __launch_bounds__(FR_T, FR_MINB) k_route_frontier(Geo g, Dev d, const int *__restrict__ env_list, FrParams P) {
    const int env = env_list[blockIdx.x];
    const int net = d.act[2 * env + 1];
    if (net <= 0) return;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int X = g.X, Y = g.Y, Z = g.Z, Xp = g.Xp;
    // ---- shared memory
    extern __shared__ __align__(16) uint32_t fsm[];
    uint32_t *sp = fsm;
    int4 *s_box = reinterpret_cast<int4 *>(sp); sp += 4 * FR_MAXPIN;
    FrSm *S = reinterpret_cast<FrSm *>(sp); sp += (sizeof(FrSm) + 15) / 16 * 4;
    uint32_t *l_base = sp; sp += 4 * (size_t)P.cap_s;      // list w: cells at l_base + 2 w cap_s, their f right behind
    uint32_t *e_pc = sp; sp += P.cap_e;                     // single-step tasks of the round: cell, distance
    uint32_t *e_d = sp; sp += P.cap_e;
    uint32_t *r_pc = sp; sp += P.cap_e;                     // ray tasks of the round: cell | direction, distance
    uint32_t *r_d = sp; sp += P.cap_e;
    int *s_xc = reinterpret_cast<int *>(sp); sp += X;
    int *s_yc = reinterpret_cast<int *>(sp); sp += Y;
    uint32_t *s_mx = sp; sp += 4 * XR_ZMAX;
    uint32_t *s_my = sp; sp += 4 * XR_ZMAX;
    uint32_t *s_mv = sp; sp += 4;
    uint32_t *s_pen = sp; sp += XR_ZMAX;
    uint32_t *s_vlen = sp; sp += XR_ZMAX;
    uint32_t *s_pref = sp; sp += XR_ZMAX;                   // 0 = rays along x, 1 = along y
    uint32_t *s_tile = sp; sp += FR_MAXTILES;               // per 16 x 16 tile: the (at most 4) boxes that can be nearest
    uint32_t *s_apcp = sp; sp += FR_MAXAP;                  // padded router index of the net's access points
    uint16_t *s_appin = reinterpret_cast<uint16_t *>(sp); sp += FR_MAXAP / 2;
    uint8_t *s_apconn = reinterpret_cast<uint8_t *>(sp); sp += FR_MAXAP / 4;
    uint8_t *s_apon = reinterpret_cast<uint8_t *>(sp); sp += FR_MAXAP / 4;
    // ---- global spill of the lists
    uint32_t *spill = d.fr_spill + (size_t)env * (8 * (size_t)P.cap_g + 4 * (size_t)P.cap_ge);
    uint32_t *farl = spill + 4 * (size_t)P.cap_g;                                     // the two far lists
    uint32_t *ge_pc = farl + 4 * (size_t)P.cap_g, *ge_d = ge_pc + P.cap_ge;          // single-step tasks
    uint32_t *gr_pc = ge_d + P.cap_ge, *gr_d = gr_pc + P.cap_ge;                      // ray tasks
    const int cap_l = P.cap_s + P.cap_g, cap_x = P.cap_e + P.cap_ge;
    auto put_l = [&](int w, int i, uint32_t pc, uint32_t f) {
        if (i < P.cap_s) { uint32_t *q = l_base + 2 * (size_t)w * P.cap_s + i; q[0] = pc; q[P.cap_s] = f; }
        else { uint32_t *q = spill + 2 * (size_t)w * P.cap_g + (i - P.cap_s); __stcg(q, pc); __stcg(q + P.cap_g, f); }
    };
    auto put_f = [&](int w, int i, uint32_t pc, uint32_t f) {
        uint32_t *q = farl + 2 * (size_t)w * P.cap_g + i; __stcg(q, pc); __stcg(q + P.cap_g, f);
    };
    auto put_e = [&](int i, uint32_t pc, uint32_t dv) {
        if (i < P.cap_e) { e_pc[i] = pc; e_d[i] = dv; }
        else { __stcg(ge_pc + (i - P.cap_e), pc); __stcg(ge_d + (i - P.cap_e), dv); }
    };
    auto put_r = [&](int i, uint32_t pc, uint32_t dv) {
        if (i < P.cap_e) { r_pc[i] = pc; r_d[i] = dv; }
        else { __stcg(gr_pc + (i - P.cap_e), pc); __stcg(gr_d + (i - P.cap_e), dv); }
    };
    // ---- tables
    for (int i = tid; i < X; i += T) s_xc[i] = g.xc[i];
    for (int i = tid; i < Y; i += T) s_yc[i] = g.yc[i];
    if (tid < 4 * XR_ZMAX) {
        const int z = tid >> 2, f = tid & 3;
        s_mx[tid] = z < Z ? g.multX[z][f] : 0u; s_my[tid] = z < Z ? g.multY[z][f] : 0u;
        if (z == 0) s_mv[f] = g.multV[f];
        if (f == 0) {
            s_pen[z] = z < Z ? g.pen[z] : 0u; s_vlen[z] = z < Z ? g.vlen[z] : 0u;
            s_pref[z] = (z < Z && g.multX[z][0] > g.multY[z][0]) ? 1u : 0u;
        }
    }
    const int ntx = (X + (1 << FR_TILE_SHIFT) - 1) >> FR_TILE_SHIFT, nty = (Y + (1 << FR_TILE_SHIFT) - 1) >> FR_TILE_SHIFT;
    for (int i = tid; i < ntx * nty; i += T) s_tile[i] = 0xFFFFFFFFu;
    if (tid == 0) { S->err = 0; S->pn = 0; S->cn = 0; S->walk_fail = 0; S->rx0 = INT_MAX; S->rx1 = 0; S->ry0 = INT_MAX; S->ry1 = 0; }
    __syncthreads();
    // ---- the net's access points
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const int ap_s = ns[net], n_ap = ns[net + 1] - ap_s;
    const size_t aoff = (size_t)env * g.max_aps + ap_s;
    const unsigned srcpin = d.net_srcpin[(size_t)env * (g.max_nets + 1) + net];
    for (int i = tid; i < n_ap; i += T) {
        const int cp = d.ap_cellp[aoff + i];
        s_apcp[i] = (uint32_t)cp;
        const unsigned pin = d.ap_pin[aoff + i];
        s_appin[i] = (uint16_t)pin;
        s_apconn[i] = pin == srcpin;
        atomicOr(d.dist64 + (size_t)env * g.cells_p + cp, FRW_OWN);       // this net's access points, for the duration of the kernel
        const int tx = (cp % Xp) >> FR_TILE_SHIFT, ty = ((cp / Xp) % Y) >> FR_TILE_SHIFT;
        atomicMin(&S->rx0, tx); atomicMax(&S->rx1, tx); atomicMin(&S->ry0, ty); atomicMax(&S->ry1, ty);
    }
    const size_t eoff = (size_t)env * g.cells_p;
    unsigned long long *dist = d.dist64 + eoff;
    // optional cost term: cells inside the net's guide boxes are marked for the duration of the kernel (a net without
    // boxes has no guide term)
    const uint32_t gcost = (uint32_t)g.guide_cost;
    int gb0 = 0, gb1 = 0;
    if (g.guide_cost > 0) { const int *gs = d.guide_start + (size_t)env * (g.max_nets + 2); gb0 = gs[net]; gb1 = gs[net + 1]; }
    const bool has_guides = gb1 > gb0;
    auto guide_marks = [&](bool set) {
        for (int b = gb0; b < gb1; b++) {
            const int *q = d.guide_box + ((size_t)env * g.guide_cap + b) * 5;
            const int x0 = max(q[0], 0), x1 = min(q[1], X - 1), y0 = max(q[2], 0), y1 = min(q[3], Y - 1), z = q[4];
            if (z < 0 || z >= Z || x1 < x0 || y1 < y0) continue;
            const int w = x1 - x0 + 1, n = w * (y1 - y0 + 1);
            for (int i = tid; i < n; i += T) {
                unsigned long long *p = dist + ((size_t)z * Y + y0 + i / w) * Xp + x0 + i % w;
                if (set) atomicOr(p, FRW_GUIDE); else atomicAnd(p, ~FRW_GUIDE);
            }
        }
    };
    guide_marks(true);
    const uint32_t *cinfo = d.cellinfo + eoff;
    const uint16_t *apn = d.apnet + eoff;
    int *path = d.path + (size_t)env * g.path_cap;
    uint32_t epoch = d.fr_epoch[env];
    int pn = 0, cn = 0;
    bool first = true;
    long long work = 0, n_rounds = 0;
    long long m_blocked = 0, m_shorted = 0, m_overflow = 0;   // congestion deltas of this thread's commits
    long long wl_acc = 0, via_acc = 0;                        // lane partial sums of warp 0
    __syncthreads();
    // tiles the heuristic tables cover: the bounding box of the net's access points + 2 tiles (cells outside take the
    // plain loop over all boxes)
    const int rtx0 = max(S->rx0 - 2, 0), rtx1 = min(S->rx1 + 2, ntx - 1), rty0 = max(S->ry0 - 2, 0), rty1 = min(S->ry1 + 2, nty - 1);
    const int rw = rtx1 - rtx0 + 1, n_reg = min(rw * (rty1 - rty0 + 1), FR_MAXTILES);
    const bool uni_x = g.uniform_x != 0, uni_y = g.uniform_y != 0;
#ifdef FR_TIMING
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, n_expanded = 0, max_open = 0, n_refill = 0;   // seed+connect, boxes+push, classify, expand, target, walk, commit
    const long long tstart = clock64();
    long long tlast = tstart;
#endif
    int n_box = 0;
    bool use_tiles = false;
    auto box_dist = [&](const int4 q, int px, int py) -> int {
        return max(max(q.x - px, px - q.y), 0) + max(max(q.z - py, py - q.w), 0);
    };
    // h: L1 track distance (DBU) to the nearest unconnected pin's access-point box
    auto hval = [&](int x, int y) -> uint32_t {
        const int px = s_xc[x], py = s_yc[y];
        int best = INT_MAX;
        const uint32_t w = use_tiles ? s_tile[(y >> FR_TILE_SHIFT) * ntx + (x >> FR_TILE_SHIFT)] : 0xFFFFFFFFu;
        if (!use_tiles) {                                   // one or two boxes
            best = box_dist(s_box[0], px, py);
            if (n_box > 1) best = min(best, box_dist(s_box[1], px, py));
        } else if (w == 0xFFFFFFFFu) {
            for (int b = 0; b < n_box; b++) best = min(best, box_dist(s_box[b], px, py));
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t b = (w >> (8 * j)) & 0xFFu;
                if (b != 0xFFu) best = min(best, box_dist(s_box[b], px, py));
            }
        }
        return (uint32_t)best;
    };
    auto lenx = [&](int xa, int xb) -> uint32_t { return uni_x ? (uint32_t)g.dx : (uint32_t)abs(s_xc[xa] - s_xc[xb]); };   // adjacent tracks
    auto leny = [&](int ya, int yb) -> uint32_t { return uni_y ? (uint32_t)g.dy : (uint32_t)abs(s_yc[ya] - s_yc[yb]); };
    // weight of the move p -> c = p + delta(dir), f = flags of c, from the shared-memory tables
    auto mw = [&](int px, int py, int pz, int dir, uint32_t f) -> uint32_t {
        uint32_t w;
        const uint32_t og = (f & CF_OG) ? gcost : 0u;
        if (dir < 2) w = lenx(px, dir == 0 ? px + 1 : px - 1) * (s_mx[4 * pz + (f & 3u)] + og) + ((f & CF_BLK) ? s_pen[pz] : 0u);
        else if (dir < 4) w = leny(py, dir == 2 ? py + 1 : py - 1) * (s_my[4 * pz + (f & 3u)] + og) + ((f & CF_BLK) ? s_pen[pz] : 0u);
        else if (dir == 4) w = s_vlen[pz] * (s_mv[f & 3u] + og) + ((f & CF_BLK) ? s_pen[pz + 1] : 0u);
        else w = s_vlen[pz - 1] * (s_mv[f & 3u] + og) + ((f & CF_BLK) ? s_pen[pz - 1] : 0u);
        return w;
    };
    // cost flags (CF_*) of a cell for this net from its word alone (the net's own access points carry FRW_OWN while it
    // is routed).  Sets own = the cell is an access point of this net.
    auto wflags = [&](unsigned long long w, size_t idx, bool &own) -> uint32_t {
        own = (w & FRW_OWN) != 0ull;
        return (uint32_t)(w & FRW_RS) | (uint32_t)((w & FRW_BLK) << 1) | (((w & (FRW_AP | FRW_OWN)) == FRW_AP) ? CF_FS : 0u) |
               ((has_guides && !(w & FRW_GUIDE)) ? CF_OG : 0u);
    };

    for (;;) {                                           // ---- one connection per trip
        epoch++;
        const uint32_t hi = (~epoch) & FRW_EPOCH_MASK;
        auto dval = [&](unsigned long long v) -> uint32_t { return (uint32_t)(v >> (FRW_SHIFT + 30)) == hi ? ((uint32_t)(v >> FRW_SHIFT) & 0x3FFFFFFFu) : XR_INF; };
        if (tid == 0) {
            S->cnt[0] = 0; S->cnt[1] = 0; S->fmin[0] = 0xFFFFFFFFu; S->fmin[1] = 0xFFFFFFFFu;
            S->nexp[0] = 0; S->nexp[1] = 0; S->nray[0] = 0; S->nray[1] = 0; S->B = XR_INF; S->nbox = 0; S->best = ~0ull; S->dmul = 1;
            S->far_n[0] = 0; S->far_n[1] = 0; S->far_min[0] = 0xFFFFFFFFu; S->far_min[1] = 0xFFFFFFFFu;
        }
        // ---- the tree so far becomes the source set of this epoch (distance 0; the flag bits of a word stay)
        if (first) {
            for (int i = tid; i < n_ap; i += T)
                if (s_apconn[i]) { const unsigned long long w = __ldcg(dist + s_apcp[i]); __stcg(dist + s_apcp[i], frw_make(hi, 0u, (uint32_t)(w & FRW_FLAGS))); }
        } else {
            for (int k = tid; k < pn; k += T) {
                const int ci = __ldcg(path + k);
                const int x = ci % X, y = (ci / X) % Y, z = ci / (X * Y);
                unsigned long long *q = dist + ((size_t)z * Y + y) * Xp + x;
                const unsigned long long w = __ldcg(q);
                __stcg(q, frw_make(hi, 0u, (uint32_t)(w & FRW_FLAGS)));
            }
        }
        __syncthreads();
        // ---- pins the tree touches are connected (any access point on the tree)
        if (!first) {
            for (int i = tid; i < n_ap; i += T)
                s_apon[i] = s_apconn[i] ? 1 : ((__ldcg(dist + s_apcp[i]) >> FRW_SHIFT) == ((unsigned long long)hi << 30));
            __syncthreads();
            for (int i = tid; i < n_ap; i += T) {
                if (s_apconn[i]) continue;
                const unsigned pin = s_appin[i];
                bool on = false;
                for (int j = i; j >= 0 && s_appin[j] == pin && !on; j--) on = s_apon[j] != 0;
                for (int j = i + 1; j < n_ap && s_appin[j] == pin && !on; j++) on = s_apon[j] != 0;
                if (on) s_apconn[i] = 2;                     // (2: the runs above only read s_apon)
            }
            __syncthreads();
        }
        bool left = false;
        for (int i = tid; i < n_ap; i += T) { if (s_apconn[i] == 2) s_apconn[i] = 1; left |= s_apconn[i] == 0; }
        if (!__syncthreads_or(left)) break;
        if (S->walk_fail || S->err) break;
        FR_TICK(0);
        // ---- boxes of the unconnected pins (heuristic)
        for (int i = tid; i < n_ap; i += T) {
            if (s_apconn[i] || (i > 0 && s_appin[i - 1] == s_appin[i])) continue;
            int x0 = INT_MAX, x1 = INT_MIN, y0 = INT_MAX, y1 = INT_MIN;
            for (int j = i; j < n_ap && s_appin[j] == s_appin[i]; j++) {
                const int cp = (int)s_apcp[j];
                const int px = s_xc[cp % Xp], py = s_yc[(cp / Xp) % Y];
                x0 = min(x0, px); x1 = max(x1, px); y0 = min(y0, py); y1 = max(y1, py);
            }
            s_box[atomicAdd(&S->nbox, 1)] = make_int4(x0, x1, y0, y1);
        }
        __syncthreads();
        n_box = S->nbox;
        // per tile of the net's region: the boxes that can be the nearest one of a cell of the tile.  A box whose
        // distance to the tile exceeds (smallest such distance + the tile's diameter) never is, so the minimum over the
        // listed boxes is exact.  Tiles with more than 4 candidates keep the plain loop.
        use_tiles = n_box > 2 && n_box < 255;
        if (use_tiles) {
            for (int t = tid; t < n_reg; t += T) {
                const int tx = rtx0 + t % rw, ty = rty0 + t / rw;
                const int ax0 = s_xc[tx << FR_TILE_SHIFT], ax1 = s_xc[min(((tx + 1) << FR_TILE_SHIFT) - 1, X - 1)];
                const int ay0 = s_yc[ty << FR_TILE_SHIFT], ay1 = s_yc[min(((ty + 1) << FR_TILE_SHIFT) - 1, Y - 1)];
                int hmin = INT_MAX;
                for (int b = 0; b < n_box; b++) {
                    const int4 q = s_box[b];
                    hmin = min(hmin, max(max(q.x - ax1, ax0 - q.y), 0) + max(max(q.z - ay1, ay0 - q.w), 0));
                }
                const int lim = hmin + (ax1 - ax0) + (ay1 - ay0);
                uint32_t word = 0xFFFFFFFFu;
                int cnt = 0;
                for (int b = 0; b < n_box; b++) {
                    const int4 q = s_box[b];
                    const int dm = max(max(q.x - ax1, ax0 - q.y), 0) + max(max(q.z - ay1, ay0 - q.w), 0);
                    if (dm <= lim) {
                        if (cnt < 4) word = (word & ~(0xFFu << (8 * cnt))) | ((uint32_t)b << (8 * cnt));
                        cnt++;
                    }
                }
                s_tile[ty * ntx + tx] = cnt > 4 ? 0xFFFFFFFFu : word;
            }
            __syncthreads();
        }
        // ---- open list 0 <- the tree cells, f = h, rays both ways
        {
            uint32_t fl = 0xFFFFFFFFu;
            const int n_src = first ? n_ap : pn;
            for (int k = tid; k < n_src; k += T) {
                int x, y, z;
                if (first) {
                    if (!s_apconn[k]) continue;
                    const int cp = (int)s_apcp[k];
                    x = cp % Xp; y = (cp / Xp) % Y; z = cp / (Xp * Y);
                } else {
                    const int ci = __ldcg(path + k);
                    x = ci % X; y = (ci / X) % Y; z = ci / (X * Y);
                }
                const uint32_t f = hval(x, y);
                const int idx = fr_agg_inc(&S->cnt[0]);
                if (idx < cap_l) put_l(0, idx, (uint32_t)(x | (y << 10) | (z << 20)) | FR_RAYS_BOTH, f); else S->err = 5;
                fl = min(fl, f);
            }
            fl = __reduce_min_sync(0xFFFFFFFFu, fl);
            if (lane == 0 && fl != 0xFFFFFFFFu) atomicMin(&S->fmin[0], fl);
        }
        FR_TICK(1);
        // ---- rounds
        // Parking: a wide search leaves behind, for every cell it expands, the entries of the steps that lead away from
        // the targets.  Their f lies above the bucket for a long time (most are dropped unexpanded when the target is
        // reached), and an open list that is classified front to back every round pays for them every round.  Once
        // the open list holds more than park_min entries, entries with f > far_thr go to a far list in global memory,
        // which is only read again when the open list has nothing below its smallest f (then far_thr moves up by
        // `band` and the entries below it come back).  The order of expansion does not change the result: the search
        // ends when no open entry -- parked ones included -- has f <= B.
        int cur = 0, par = 0, fc = 0;
        uint32_t far_thr = 0xFFFFFFFFu;
        for (;;) {
            __syncthreads();                              // (A) the lists, counters and B of the previous round are published
            FR_TICK(3);
            const int n = min(S->cnt[cur], cap_l);
            const uint32_t fm = S->fmin[cur], Bv = S->B;
            int fn = 0;
            uint32_t fmn = 0xFFFFFFFFu;
            if (WIDE && far_thr != 0xFFFFFFFFu) {         // (the classification below appends to the far list: everyone reads first)
                fn = min(S->far_n[fc], P.cap_g); fmn = S->far_min[fc];
                __syncthreads();
            }
            if (WIDE && fn > 0 && fmn <= Bv && (n == 0 || fm >= fmn)) {
                // ---- the far list comes back: entries up to the new far_thr join the open list, the rest is compacted into
                // the other far list (stale entries and entries above B are dropped on the way)
                const uint32_t nthr = min(fm, fmn) + P.band;
                const int fo = fc ^ 1;
                uint32_t fl_n = 0xFFFFFFFFu, fl_f = 0xFFFFFFFFu;
                const uint32_t *fq = farl + 2 * (size_t)fc * P.cap_g;
                for (int i0 = 0; i0 < fn; i0 += 4 * T) {
                    uint32_t pc[4], f[4];
                    unsigned long long wv[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int i = i0 + k * T + tid;
                        f[k] = 0xFFFFFFFFu; pc[k] = 0;
                        if (i < fn) { pc[k] = __ldcg(fq + i); f[k] = __ldcg(fq + P.cap_g + i); }
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        wv[k] = 0ull;
                        if (f[k] <= Bv) wv[k] = __ldcg(dist + ((size_t)((pc[k] >> 20) & 15) * Y + ((pc[k] >> 10) & 1023)) * Xp + (pc[k] & 1023));
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (f[k] > Bv) continue;
                        if (dval(wv[k]) < f[k] - hval(pc[k] & 1023, (pc[k] >> 10) & 1023)) continue;
                        if (f[k] <= nthr) {
                            const int idx = fr_agg_inc(&S->cnt[cur]);
                            if (idx < cap_l) put_l(cur, idx, pc[k], f[k]); else S->err = 5;
                            fl_n = min(fl_n, f[k]);
                        } else {
                            const int idx = fr_agg_inc(&S->far_n[fo]);
                            if (idx < P.cap_g) put_f(fo, idx, pc[k], f[k]); else S->err = 5;
                            fl_f = min(fl_f, f[k]);
                        }
                    }
                }
                fl_n = __reduce_min_sync(0xFFFFFFFFu, fl_n); fl_f = __reduce_min_sync(0xFFFFFFFFu, fl_f);
                if (lane == 0 && fl_n != 0xFFFFFFFFu) atomicMin(&S->fmin[cur], fl_n);
                if (lane == 0 && fl_f != 0xFFFFFFFFu) atomicMin(&S->far_min[fo], fl_f);
                __syncthreads();
                if (tid == 0) { S->far_n[fc] = 0; S->far_min[fc] = 0xFFFFFFFFu; }
                fc = fo; far_thr = nthr;
                if (S->err) break;
#ifdef FR_TIMING
                { const long long t__ = clock64(); ph[7] += t__ - tlast; tlast = t__; n_refill++; }
#endif
                continue;
            }
            if (n == 0 || fm > Bv) break;
            n_rounds++;
            const uint32_t thr = fm + P.delta * (uint32_t)S->dmul;
            if (WIDE && far_thr == 0xFFFFFFFFu && P.park_min > 0 && n > P.park_min) far_thr = fm + P.band;
            // list pressure (a quarter of the capacity in use): drop the stale entries of a cell here and push a cell only
            // when its word really went down, so that the list holds at most one live entry per cell and lowering
            const bool tight = n + fn > cap_l / 4;
            const int nxt = cur ^ 1;
            uint32_t fl = 0xFFFFFFFFu, flf = 0xFFFFFFFFu;
            // classify: expand now (one single-step task + a ray task per ray bit) / keep for later / drop (f > B)
            auto classify = [&](const uint32_t pc, const uint32_t f) {
                if (f > Bv) return;
                if (tight) {
                    const int ex = pc & 1023, ey = (pc >> 10) & 1023, ez = (pc >> 20) & 15;
                    if (dval(__ldcg(dist + ((size_t)ez * Y + ey) * Xp + ex)) < f - hval(ex, ey)) return;
                }
                if (f <= thr) {
                    const uint32_t d0 = f - hval(pc & 1023, (pc >> 10) & 1023);
                    const int idx = fr_agg_inc(&S->nexp[par]);
                    if (idx < cap_x) put_e(idx, pc & 0x00FFFFFFu, d0); else S->err = 5;
                    const int nr = ((pc >> 24) & 1u) + ((pc >> 25) & 1u);
                    if (nr) {
                        int q = atomicAdd(&S->nray[par], nr);
                        if (pc & FR_RAY_POS) { if (q < cap_x) put_r(q, pc & 0x00FFFFFFu, d0); else S->err = 5; q++; }
                        if (pc & FR_RAY_NEG) { if (q < cap_x) put_r(q, (pc & 0x00FFFFFFu) | FR_RAY_NEG, d0); else S->err = 5; }
                    }
                } else if (!WIDE || f <= far_thr) {
                    const int idx = fr_agg_inc(&S->cnt[nxt]);
                    if (idx < cap_l) put_l(nxt, idx, pc, f); else S->err = 5;
                    fl = min(fl, f);
                } else {
                    const int idx = fr_agg_inc(&S->far_n[fc]);
                    if (idx < P.cap_g) put_f(fc, idx, pc, f); else S->err = 5;
                    flf = min(flf, f);
                }
            };
            if constexpr (WIDE) {
                for (int i0 = 0; i0 < n; i0 += 4 * T) {     // four entries per thread in flight: the spilled part is an L2 round trip each
                    uint32_t pc4[4], f4[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int i = i0 + k * T + tid;
                        pc4[k] = 0; f4[k] = 0xFFFFFFFFu;
                        if (i < n) {
                            if (i < P.cap_s) { const uint32_t *q = l_base + 2 * (size_t)cur * P.cap_s + i; pc4[k] = q[0]; f4[k] = q[P.cap_s]; }
                            else { const uint32_t *q = spill + 2 * (size_t)cur * P.cap_g + (i - P.cap_s); pc4[k] = __ldcg(q); f4[k] = __ldcg(q + P.cap_g); }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) classify(pc4[k], f4[k]);
                }
            } else {
                for (int i = tid; i < n; i += T) {
                    uint32_t pc, f;
                    if (i < P.cap_s) { const uint32_t *q = l_base + 2 * (size_t)cur * P.cap_s + i; pc = q[0]; f = q[P.cap_s]; }
                    else { const uint32_t *q = spill + 2 * (size_t)cur * P.cap_g + (i - P.cap_s); pc = __ldcg(q); f = __ldcg(q + P.cap_g); }
                    classify(pc, f);
                }
            }
            __syncthreads();                              // (B)
            FR_TICK(2);
            const int nE = min(S->nexp[par], cap_x), nR = min(S->nray[par], cap_x);
#ifdef FR_TIMING
            n_expanded += nE; max_open = max(max_open, (long long)n);
#endif
            if (tid == 0) {
                S->cnt[cur] = 0; S->fmin[cur] = 0xFFFFFFFFu; S->nexp[par ^ 1] = 0; S->nray[par ^ 1] = 0;
                // bucket width: a round costs the same (one L2 round trip + two barriers) whether it expands ten
                // entries or a thousand -- widen the bucket while rounds are nearly empty, narrow it when they are wide
                if (nE < T / 16) S->dmul = min(S->dmul * 2, P.dmax); else if (nE > 2 * T) S->dmul = max(S->dmul / 2, 1);
            }
            // a lowered cell that is an access point of this net: a target if its pin is unconnected
            auto target_check = [&](size_t idx, uint32_t nd) {
                for (int i = 0; i < n_ap; i++)
                    if (s_apcp[i] == (uint32_t)idx) { if (!s_apconn[i]) atomicMin(&S->B, nd); break; }
            };
            // Push a lowered cell (warp-aggregated append to the next list).  Called by all 32 lanes.
            auto push = [&](bool keep, uint32_t pcv, uint32_t fv) {
                const bool park = WIDE && keep && fv > far_thr;
                keep = keep && !park;
                const unsigned mk = __ballot_sync(0xFFFFFFFFu, keep), mp = WIDE ? __ballot_sync(0xFFFFFFFFu, park) : 0u;
                if ((mk | mp) == 0u) return;
                int base = 0, basep = 0;
                if (lane == 0 && mk) base = atomicAdd(&S->cnt[nxt], __popc(mk));
                if (WIDE && lane == 0 && mp) basep = atomicAdd(&S->far_n[fc], __popc(mp));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (WIDE) basep = __shfl_sync(0xFFFFFFFFu, basep, 0);
                if (keep) {
                    const int q = base + __popc(mk & ((1u << lane) - 1u));
                    if (q < cap_l) put_l(nxt, q, pcv, fv); else S->err = 5;
                    fl = min(fl, fv);
                }
                if (park) {
                    const int q = basep + __popc(mp & ((1u << lane) - 1u));
                    if (q < P.cap_g) put_f(fc, q, pcv, fv); else S->err = 5;
                    flf = min(flf, fv);
                }
            };
            // ---- ray tasks, one LANE per cell: 8 lanes relax up to `ray` cells along the layer's preferred direction.
            // Each lane loads its cell's word, a segmented prefix sum of the edge weights gives the tentative
            // distances, the ray ends at the first cell it does not lower.  Only the last cell of a full-length ray
            // shoots on (same direction); the others were relaxed onwards by this very ray, and backwards lies the
            // smaller distance they came from.
            // Ray lanes first, single-step lanes behind them (from a warp boundary on), dealt to the warps round robin.
            // A lane's task of the next trip is fetched before the current one is worked on: tasks beyond cap_e live in
            // global memory, and the fetch would otherwise sit in front of the cell loads on the critical path.
            const int R32 = (FR_RAY * nR + 31) & ~31, n_lanes = R32 + 4 * nE;
            auto fetch = [&](int w0, uint32_t &pc, uint32_t &d0) {
                if (w0 >= n_lanes) return;
                if (w0 < R32) {
                    const int w = w0 + lane, t = w < FR_RAY * nR ? w / FR_RAY : 0;
                    if (t < P.cap_e) { pc = r_pc[t]; d0 = r_d[t]; }
                    else { pc = __ldcg(gr_pc + (t - P.cap_e)); d0 = __ldcg(gr_d + (t - P.cap_e)); }
                } else {
                    const int w = w0 - R32 + lane, t = w < 4 * nE ? w >> 2 : 0;
                    if (t < P.cap_e) { pc = e_pc[t]; d0 = e_d[t]; }
                    else { pc = __ldcg(ge_pc + (t - P.cap_e)); d0 = __ldcg(ge_d + (t - P.cap_e)); }
                }
            };
            uint32_t pc_n = 0, d0_n = 0;
            if (WIDE) fetch(warp * 32, pc_n, d0_n);
            for (int w0 = warp * 32; w0 < n_lanes; w0 += T) {
              uint32_t pc = pc_n, d0 = d0_n;
              if (WIDE) fetch(w0 + T, pc_n, d0_n);
              if (w0 < R32) {
                const int w = w0 + lane;
                const bool act = w < FR_RAY * nR;
                const int k = w & (FR_RAY - 1);
                if constexpr (!WIDE) {
                    const int t = act ? w / FR_RAY : 0;
                    if (t < P.cap_e) { pc = r_pc[t]; d0 = r_d[t]; }
                    else { pc = __ldcg(gr_pc + (t - P.cap_e)); d0 = __ldcg(gr_d + (t - P.cap_e)); }
                }
                const int x = pc & 1023, y = (pc >> 10) & 1023, z = (pc >> 20) & 15;
                const int sgn = (pc & FR_RAY_NEG) ? -1 : 1;
                const size_t cp0 = ((size_t)z * Y + y) * Xp + x;
                const bool alongx = s_pref[z] == 0;
                const int room = alongx ? (sgn > 0 ? X - 1 - x : x) : (sgn > 0 ? Y - 1 - y : y);
                const int nr = min(P.ray, room);
                const bool valid = act && k < nr;
                const size_t idx = cp0 + (long long)(k + 1) * (alongx ? sgn : sgn * Xp);
                const int xv = alongx ? x + sgn * (k + 1) : x, yv = alongx ? y : y + sgn * (k + 1);
                const unsigned long long wv = valid ? __ldcg(dist + idx) : 0ull;
                const unsigned long long own = act ? __ldcg(dist + cp0) : 0ull;
                const bool fresh = valid && dval(own) >= d0;      // (a stale entry: the cell was lowered after it was pushed)
                bool isown = false;
                uint32_t wgt = 0;
                if (valid) {
                    const uint32_t f = wflags(wv, idx, isown);
                    const uint32_t len = alongx ? lenx(xv, xv - sgn) : leny(yv, yv - sgn);
                    wgt = len * ((alongx ? s_mx : s_my)[4 * z + (f & 3u)] + ((f & CF_OG) ? gcost : 0u)) + ((f & CF_BLK) ? s_pen[z] : 0u);
                }
#pragma unroll
                for (int off = 1; off < FR_RAY; off <<= 1) {          // inclusive prefix sum inside the 8-lane group
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, wgt, off, FR_RAY);
                    if (k >= off) wgt += o;
                }
                const uint32_t nd = d0 + wgt;
                const uint32_t Bnow = *(volatile uint32_t *)&S->B;
                const bool okk = fresh && nd <= Bnow && nd < dval(wv);
                const unsigned bits = (__ballot_sync(0xFFFFFFFFu, okk) >> (lane & ~(FR_RAY - 1))) & ((1u << FR_RAY) - 1u);
                const int m = __ffs(~bits) - 1;                       // cells of this ray that are lowered (a prefix)
                const bool low = valid && k < m;
                work += fresh ? 1 : 0;
                bool won = low;
                if (low) {
                    const unsigned long long key = frw_make(hi, nd, (uint32_t)(wv & FRW_FLAGS));
                    if (tight) won = atomicMin(dist + idx, key) > key; else atomicMin(dist + idx, key);
                    if (isown) target_check(idx, nd);
                }
                __syncwarp();
                const uint32_t B2 = *(volatile uint32_t *)&S->B;
                uint32_t fv = 0;
                if (won) fv = nd + hval(xv, yv);
                const uint32_t on = (k == m - 1 && m == P.ray) ? (sgn > 0 ? FR_RAY_POS : FR_RAY_NEG) : 0u;
                push(won && fv <= B2, (uint32_t)(xv | (yv << 10) | (z << 20)) | on, fv);
              } else {
                // ---- single-step tasks, one lane per cell: the two wrong-way neighbours and the two vias of every expanded cell
                const int w = w0 - R32 + lane;
                const bool act = w < 4 * nE;
                const int j = w & 3;
                if constexpr (!WIDE) {
                    const int t = act ? w >> 2 : 0;
                    if (t < P.cap_e) { pc = e_pc[t]; d0 = e_d[t]; }
                    else { pc = __ldcg(ge_pc + (t - P.cap_e)); d0 = __ldcg(ge_d + (t - P.cap_e)); }
                }
                const int x = pc & 1023, y = (pc >> 10) & 1023, z = (pc >> 20) & 15;
                const bool alongx = s_pref[z] == 0;
                const int sgn = (j & 1) ? -1 : 1;
                const int xv = x + ((j < 2 && !alongx) ? sgn : 0), yv = y + ((j < 2 && alongx) ? sgn : 0), zv = z + (j >= 2 ? sgn : 0);
                const bool valid = act && xv >= 0 && xv < X && yv >= 0 && yv < Y && zv >= 0 && zv < Z;
                const size_t idx = ((size_t)zv * Y + yv) * Xp + xv;
                const unsigned long long wv = valid ? __ldcg(dist + idx) : 0ull;
                const unsigned long long own = act ? __ldcg(dist + ((size_t)z * Y + y) * Xp + x) : 0ull;
                const bool fresh = valid && dval(own) >= d0;
                bool isown = false;
                uint32_t nd = 0;
                if (valid) {
                    const uint32_t f = wflags(wv, idx, isown);
                    uint32_t wgt;
                    const uint32_t og = (f & CF_OG) ? gcost : 0u;
                    if (j < 2) wgt = (alongx ? leny(yv, y) : lenx(xv, x)) * ((alongx ? s_my : s_mx)[4 * z + (f & 3u)] + og);
                    else wgt = s_vlen[j == 2 ? z : z - 1] * (s_mv[f & 3u] + og);
                    nd = d0 + wgt + ((f & CF_BLK) ? s_pen[zv] : 0u);
                }
                const uint32_t Bnow = *(volatile uint32_t *)&S->B;
                const bool low = fresh && nd <= Bnow && nd < dval(wv);
                work += fresh ? 1 : 0;
                bool won = low;
                if (low) {
                    const unsigned long long key = frw_make(hi, nd, (uint32_t)(wv & FRW_FLAGS));
                    if (tight) won = atomicMin(dist + idx, key) > key; else atomicMin(dist + idx, key);
                    if (isown) target_check(idx, nd);
                }
                __syncwarp();
                const uint32_t B2 = *(volatile uint32_t *)&S->B;
                uint32_t fv = 0;
                if (won) fv = nd + hval(xv, yv);
                push(won && fv <= B2, (uint32_t)(xv | (yv << 10) | (zv << 20)) | FR_RAYS_BOTH, fv);
              }
            }
            fl = __reduce_min_sync(0xFFFFFFFFu, fl); flf = __reduce_min_sync(0xFFFFFFFFu, flf);
            if (lane == 0 && fl != 0xFFFFFFFFu) atomicMin(&S->fmin[nxt], fl);
            if (WIDE && lane == 0 && flf != 0xFFFFFFFFu) atomicMin(&S->far_min[fc], flf);
            cur = nxt; par ^= 1;
        }
        if (S->err) break;
        // ---- target = argmin (distance, padded cell index) over the access points of the unconnected pins
        {
            unsigned long long best = ~0ull;
            for (int i = tid; i < n_ap; i += T) {
                if (s_apconn[i]) continue;
                const uint32_t dv = dval(__ldcg(dist + s_apcp[i]));
                const unsigned long long key = ((unsigned long long)dv << 32) | s_apcp[i];
                best = key < best ? key : best;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, best, off);
                best = o < best ? o : best;
            }
            if (lane == 0 && best != ~0ull) atomicMin(&S->best, best);
        }
        __syncthreads();
        FR_TICK(4);
        const unsigned long long best = S->best;
        const uint32_t Bfin = (uint32_t)(best >> 32);
        if (best == ~0ull || Bfin >= XR_INF) { if (tid == 0) S->err = 1; __syncthreads(); break; }
        // ---- canonical backtrace by warp 0: predecessor order [last, +x, -x, +y, -y, +z, -z].  One round trip probes all
        // six directions several cells deep (a lane checks the step from c - k delta to c - (k+1) delta); a run that is
        // still going at full depth continues 32 cells per round trip along its direction.  The walk reads
        // the field of this connection only; its cells join the tree afterwards.
        const int pn0 = pn;
        if (warp == 0) {
            const int cpb = (int)(best & 0xFFFFFFFFu);
            int cx = cpb % Xp, cy = (cpb / Xp) % Y, cz = cpb / (Xp * Y);
            int last = -1, fail = 0, q = pn;
            uint32_t dc = Bfin;
            auto inb = [&](int x, int y, int z) { return x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z; };
            auto W = [&](int x, int y, int z) -> unsigned long long { return __ldcg(dist + ((size_t)z * Y + y) * Xp + x); };
            // one step of the walk checked by one lane: is b = a - delta(dir) a valid predecessor of a?  Returns d(b).
            auto step_ok = [&](int ax, int ay, int az, int dir, bool &ok) -> uint32_t {
                int ddx, ddy, ddz; dir_delta(dir, ddx, ddy, ddz);
                const int bx = ax - ddx, by = ay - ddy, bz = az - ddz;
                ok = inb(ax, ay, az) && inb(bx, by, bz);
                uint32_t db = XR_INF;
                if (ok) {
                    const unsigned long long wa = W(ax, ay, az);
                    db = dval(W(bx, by, bz));
                    const uint32_t da = dval(wa);
                    bool isown;
                    const uint32_t fa = wflags(wa, ((size_t)az * Y + ay) * Xp + ax, isown);
                    ok = da != 0 && da < XR_INF && db < XR_INF && db + mw(bx, by, bz, dir, fa) == da;
                }
                return db;
            };
            auto record = [&](int ax, int ay, int az, int dir, int slot) {
                int ddx, ddy, ddz; dir_delta(dir, ddx, ddy, ddz);
                if (slot < g.path_cap) __stcg(path + slot, (az * Y + ay) * X + ax);
                if (dir >= 4) via_acc += 1;
                else if (dir < 2) wl_acc += abs(s_xc[ax] - s_xc[ax - ddx]);
                else wl_acc += abs(s_yc[ay] - s_yc[ay - ddy]);
            };
            while (dc != 0) {
                __syncwarp();
                // one round trip probes all six directions: the two preferred directions of the cell's layer 11 steps deep
                // (lanes 0-10, 11-21), the other four 2 steps deep (lanes 22-29) -- runs are long along the preferred
                // direction and short across it
                const int p0 = s_pref[cz] == 0 ? 0 : 2;
                auto base_of = [&](int dd) { return dd == p0 ? 0 : dd == p0 + 1 ? 11 : 22 + 2 * (dd >= 4 ? dd - 2 : (dd & 1)); };
                auto depth_of = [&](int dd) { return (dd == p0 || dd == p0 + 1) ? 11 : 2; };
                int dir, k;
                if (lane < 22) { dir = p0 + (lane >= 11 ? 1 : 0); k = lane - (lane >= 11 ? 11 : 0); }
                else { const int o = (lane - 22) >> 1; dir = o < 2 ? (p0 == 0 ? 2 + o : o) : 2 + o; k = (lane - 22) & 1; }
                int ddx = 0, ddy = 0, ddz = 0;
                bool ok = false;
                uint32_t db = XR_INF;
                if (lane < 30) {
                    dir_delta(dir, ddx, ddy, ddz);
                    db = step_ok(cx - k * ddx, cy - k * ddy, cz - k * ddz, dir, ok);
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                int pick = -1;
                if (last >= 0 && ((m >> base_of(last)) & 1u)) pick = last;
                else {
#pragma unroll
                    for (int dd = 5; dd >= 0; dd--) if ((m >> base_of(dd)) & 1u) pick = dd;
                }
                if (pick < 0) { fail = 3; break; }
                const int pbase = base_of(pick), pdepth = depth_of(pick);
                const unsigned bits = (m >> pbase) & ((1u << pdepth) - 1u);
                const int run = bits == ((1u << pdepth) - 1u) ? pdepth : (__ffs(~bits) - 1);
                if (dir == pick && k < run && lane < 30) record(cx - k * ddx, cy - k * ddy, cz - k * ddz, dir, q + k);
                q += run;
                dc = __shfl_sync(0xFFFFFFFFu, db, pbase + run - 1);
                int pdx, pdy, pdz; dir_delta(pick, pdx, pdy, pdz);
                cx -= run * pdx; cy -= run * pdy; cz -= run * pdz;
                last = pick;
                // a straight run that is still going: 32 cells per round trip
                while (dc != 0 && run == pdepth) {
                    __syncwarp();
                    bool ok2 = false;
                    const uint32_t db2 = step_ok(cx - lane * pdx, cy - lane * pdy, cz - lane * pdz, last, ok2);
                    const unsigned m2 = __ballot_sync(0xFFFFFFFFu, ok2);
                    const int run2 = (m2 == 0xFFFFFFFFu) ? 32 : (__ffs(~m2) - 1);
                    if (run2 == 0) break;
                    if (lane < run2) record(cx - lane * pdx, cy - lane * pdy, cz - lane * pdz, last, q + lane);
                    q += run2;
                    dc = __shfl_sync(0xFFFFFFFFu, db2, run2 - 1);
                    cx -= run2 * pdx; cy -= run2 * pdy; cz -= run2 * pdz;
                    if (run2 < 32) break;
                }
            }
            if (!fail) {                                  // the tree cell the walk ended on closes the recorded path
                if (lane == 0 && q < g.path_cap) __stcg(path + q, (cz * Y + cy) * X + cx);
                q += 1;
            }
            if (q > g.path_cap) fail = 4;
            if (lane == 0) {
                if (cn < g.conn_cap) {
                    d.conn_cost[(size_t)env * g.conn_cap + cn] = Bfin;
                    d.conn_off[(size_t)env * (g.conn_cap + 1) + cn + 1] = q;
                }
                S->pn = q; S->cn = cn + 1; S->walk_fail = fail;
            }
        }
        __syncthreads();
        FR_TICK(5);
        pn = S->pn; cn = S->cn;
        if (S->walk_fail) break;
        // ---- commit the cells that are new to the tree: all on the first connection, all but the last afterwards
        for (int k = pn0 + tid; k < (first ? pn : pn - 1); k += T) {
            const int ci = __ldcg(path + k);
            const int x = ci % X, y = (ci / X) % Y, z = ci / (X * Y);
            const size_t c = ((size_t)z * Y + y) * Xp + x;
            uint32_t v = cinfo[c];
            const uint32_t us = (v & CI_USAGE_MASK) >> CI_USAGE_SHIFT;
            if (us == 0u) {
                m_blocked += (v & CI_BLOCK) ? 1 : 0;
                m_shorted += ((v & CI_ISAP) && apn[c] != (uint16_t)net) ? 1 : 0;
            } else if (us == 1u) {
                m_shorted += ((v & CI_ISAP) && apn[c] != (v & CI_OWNER_MASK)) ? 0 : 1;
                m_overflow += 1;
            } else if (us < 255u) m_overflow += 1;
            if (us < 255u) v += 1u << CI_USAGE_SHIFT;
            if ((v & CI_OWNER_MASK) == 0u) v |= (uint32_t)net;
            d.cellinfo[eoff + c] = v;
            atomicOr(dist + c, FRW_RS);                      // the cell word: a wire covers the cell from now on
            const size_t oo = ((size_t)x * Y + y) * Z + z;
            d.obst_obs[(size_t)env * g.cells_o + oo] = 1;
            d.obs[(size_t)env * g.obs_stride + oo] = 1.f;   // channel 0 of the observation, updated in place
        }
        first = false;
        __syncthreads();
        FR_TICK(6);
    }
    // ---- write back
    __syncthreads();
    for (int i = tid; i < n_ap; i += T) atomicAnd(dist + s_apcp[i], ~FRW_OWN);
    guide_marks(false);
    if (g.halo > 0 && !S->walk_fail && !S->err) {
        // optional cost term: the spacing halo of this net's wires counts as route shape for the nets routed later
        const int hw = 2 * g.halo + 1;
        for (int i = tid; i < pn * hw * hw; i += T) {
            const int ci = __ldcg(path + i / (hw * hw)), r = i % (hw * hw);
            const int x = ci % X + r % hw - g.halo, y = (ci / X) % Y + r / hw - g.halo, z = ci / (X * Y);
            if (x >= 0 && x < X && y >= 0 && y < Y) atomicOr(dist + ((size_t)z * Y + y) * Xp + x, FRW_RS);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        work += __shfl_xor_sync(0xFFFFFFFFu, work, off);
        m_blocked += __shfl_xor_sync(0xFFFFFFFFu, m_blocked, off);
        m_shorted += __shfl_xor_sync(0xFFFFFFFFu, m_shorted, off);
        m_overflow += __shfl_xor_sync(0xFFFFFFFFu, m_overflow, off);
        wl_acc += __shfl_xor_sync(0xFFFFFFFFu, wl_acc, off);
        via_acc += __shfl_xor_sync(0xFFFFFFFFu, via_acc, off);
    }
    if (lane == 0) {
        unsigned long long *es = reinterpret_cast<unsigned long long *>(d.envstat + 8 * (size_t)env);
        if (work) atomicAdd(es + 7, (unsigned long long)work);
        unsigned long long *mc = reinterpret_cast<unsigned long long *>(d.minc + 4 * (size_t)env);
        if (m_blocked) atomicAdd(mc + 0, (unsigned long long)m_blocked);
        if (m_shorted) atomicAdd(mc + 1, (unsigned long long)m_shorted);
        if (m_overflow) atomicAdd(mc + 2, (unsigned long long)m_overflow);
    }
#ifdef FR_TIMING
    if (tid == 0 && d.dbg) {
        const unsigned long long tot = (unsigned long long)(clock64() - tstart);
        atomicAdd(&d.dbg[0], tot); atomicMax(&d.dbg[1], tot);
        for (int k = 0; k < 7; k++) atomicAdd(&d.dbg[2 + k], (unsigned long long)ph[k]);
        atomicAdd(&d.dbg[9], (unsigned long long)n_rounds); atomicAdd(&d.dbg[10], (unsigned long long)cn);
        atomicAdd(&d.dbg[11], (unsigned long long)n_expanded); atomicMax(&d.dbg[12], (unsigned long long)n_rounds);
        atomicAdd(&d.dbg[13], 1ull);
        if (blockIdx.x == 0) { atomicAdd(&d.dbg[14], tot); atomicAdd(&d.dbg[15], (unsigned long long)n_rounds); }
        unsigned long long *rec = d.dbg + 16 + 8 * (size_t)env;
        rec[0] = tot; rec[1] = (unsigned long long)n_rounds; rec[2] = (unsigned long long)n_expanded; rec[3] = (unsigned long long)cn;
        rec[4] = (unsigned long long)ph[3]; rec[5] = (unsigned long long)ph[2]; rec[6] = (unsigned long long)ph[7] | ((unsigned long long)n_refill << 40); rec[7] = (unsigned long long)max_open;
    }
#endif
    if (tid == 0) {
        d.fr_epoch[env] = epoch;
        d.path_n[env] = pn; d.conn_n[env] = cn;
        d.wlvia[2 * env] += wl_acc; d.wlvia[2 * env + 1] += via_acc;
        d.envstat[8 * (size_t)env + 2] += n_rounds;
        d.envstat[8 * (size_t)env + 3] += cn;
        const int code = S->walk_fail ? S->walk_fail : S->err;
        if (code) d.flags[1] = code;
    }
}

size_t xr_frontier_smem(const Geo &g, const FrParams &P) {
    size_t w = 4 * FR_MAXPIN + (sizeof(FrSm) + 15) / 16 * 4 + 4 * (size_t)P.cap_s + 4 * (size_t)P.cap_e + g.X + g.Y +
               4 * XR_ZMAX * 2 + 4 + 3 * XR_ZMAX + 4096 + FR_MAXAP + FR_MAXAP / 2 + FR_MAXAP / 4 + FR_MAXAP / 4;
    return w * 4;
}
size_t xr_frontier_spill_words(const FrParams &P) { return 8 * (size_t)P.cap_g + 4 * (size_t)P.cap_ge; }

cudaError_t xr_frontier_init(int smem_cap) {
    cudaError_t e = cudaFuncSetAttribute(k_route_frontier<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_route_frontier<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_cap);
}

cudaError_t xr_frontier_launch(const Geo &g, const Dev &d, const int *env_list, int n_envs, const FrParams &P,
                               int threads, cudaStream_t st) {
    if (P.park_min > 0) k_route_frontier<true><<<n_envs, threads, xr_frontier_smem(g, P), st>>>(g, d, env_list, P);
    else k_route_frontier<false><<<n_envs, threads, xr_frontier_smem(g, P), st>>>(g, d, env_list, P);
    return cudaGetLastError();
}
