// xr_kernels_env.cuh -- observation build, congestion/reward reduction, reset and
// per-step bookkeeping kernels (all HBM-bound, coalesced, 16-byte accesses).
#pragma once
#include "xr_common.cuh"

// ------------------------------------------------------------------ observation
// Replaces baseline/build_3Dgrid.py:94-188,224-270 (getObstacleGrid, getNetGrid,
// getNetOrderChannel, _build_3Dgrid, the t.cat copy).  One pass: every CTA owns a
// contiguous chunk of the environment's [2+7n][cells] float block, streams it out
// with 16-byte stores (channel 0 from the obstacle bytes, channel 1 from the rank
// list, net channels zero) and, after a block barrier, patches the access points
// that fall inside its chunk.  Algorithmic bytes: 4*(2+7n)*cells written per env.
// Which environments a step-epilogue pass covers.  pass >= 0: the environments of post-route
// group `pass` whose route is complete (phase 0) and that no earlier pass finalised;
// pass < 0: everything still pending (after the full-grid route loop).  Tag = pass + 2 / 1.
__device__ __forceinline__ bool pass_selects(const Dev &d, int env, int pass) {
    if (d.fin[env] != 0) return false;
    if (pass < 0) return true;
    // (a net on the full-grid path is pumped on its own stream beside the group passes: it may complete at any moment,
    // so it always waits for the last pass)
    if (d.grp[env] != pass || (d.mode[env] == 0 && d.act[2 * env + 1] != 0)) return false;
    return d.phase[env] == 0;
}
__device__ __forceinline__ int pass_tag(int pass) { return pass < 0 ? 1 : pass + 2; }

#define OBS_THREADS 256
#define OBS_F4_PER_THREAD 16
#define OBS_CHUNK (OBS_THREADS * OBS_F4_PER_THREAD * 4)   // floats per CTA (64 KB)

__device__ __forceinline__ void st_stream_f4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(OBS_THREADS) k_obs(Geo g, Dev d, int tag) {
    const int env = blockIdx.y;
    if (!d.obs_do[env] || (tag > 0 && d.fin[env] != tag) || (tag == 0 && !d.obs_full[env])) return;   // tag -1: all flagged
    const int n_rem = d.n_remaining[env];
    const int n = n_rem < g.obs_max_nets ? n_rem : g.obs_max_nets;
    const long long cells = g.cells;
    const long long total = (2ll + 7ll * n) * cells;
    const long long lo = (long long)blockIdx.x * OBS_CHUNK;
    if (lo >= total) return;
    const long long hi = (lo + OBS_CHUNK < total) ? lo + OBS_CHUNK : total;
    float *out = d.obs + (size_t)env * g.obs_stride;
    const uint8_t *ob = d.obst_obs + (size_t)env * g.cells_o;
    const int *rank = d.rank_net + (size_t)env * g.max_nets;
    const long long special_end = 2 * cells;
    // ---- stream the chunk
#pragma unroll 4
    for (int k = 0; k < OBS_F4_PER_THREAD; k++) {
        const long long p = lo + ((long long)k * OBS_THREADS + threadIdx.x) * 4;
        if (p >= hi) break;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < special_end) {
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const long long q = p + j;
                float val = 0.f;
                if (q < cells) val = ob[q] ? 1.f : 0.f;
                else if (q < special_end) { const long long idx = q - cells; val = idx < n_rem ? (float)rank[idx] : 0.f; }
                e[j] = val;
            }
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        st_stream_f4(out + p, v);
    }
    if (hi <= special_end || n == 0) return;
    __syncthreads();
    // ---- patch the access points of the net blocks that overlap [lo, hi)
    const long long chlo = lo / cells, chhi = (hi - 1) / cells;
    int r0 = chlo >= 2 ? (int)((chlo - 2) / 7) : 0;
    int r1 = (int)((chhi - 2) / 7);
    if (r1 > n - 1) r1 = n - 1;
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const size_t aoff = (size_t)env * g.max_aps;
    for (int r = r0; r <= r1; r++) {
        const int net = rank[r];
        const long long base = (2ll + 7ll * r) * cells;
        const int s = ns[net], t = ns[net + 1];
        for (int i = s + threadIdx.x; i < t; i += OBS_THREADS) {
            const long long p0 = base + d.ap_obsoff[aoff + i];
            if (p0 >= lo && p0 < hi) out[p0] = 1.f;
            if (d.ap_adj[aoff + i]) {
#pragma unroll
                for (int j = 1; j <= 6; j++) {
                    const long long pj = p0 + j * cells;
                    if (pj >= lo && pj < hi) out[pj] = 1.f;
                }
            }
        }
    }
}

// ------------------------------------------------------- incremental observation
// The observation buffer persists between steps and, outside channels 0 and 1, is zero
// everywhere except at the access points of the remaining nets (invariant, established by
// the first full build).  Routing net a removes its 7-channel block and shifts every later
// block down by one (channel compaction) -- which only moves the few
// access-point ones: clear them at the old block, set them at the new one.  Channel 0 is
// patched by commit_cell, channel 1 is rewritten here.  One CTA per environment.
__device__ __forceinline__ void obs_mark_net(const Geo &g, const Dev &d, int env, int net, int block, float val) {
    if (block >= g.obs_max_nets) return;
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const size_t aoff = (size_t)env * g.max_aps;
    float *out = d.obs + (size_t)env * g.obs_stride + (2ll + 7ll * block) * g.cells;
    for (int i = ns[net] + threadIdx.x; i < ns[net + 1]; i += blockDim.x) {
        const int off = d.ap_obsoff[aoff + i];
        out[off] = val;
        if (d.ap_adj[aoff + i]) {
#pragma unroll
            for (int j = 1; j <= 6; j++) out[(size_t)j * g.cells + off] = val;
        }
    }
}

__global__ void __launch_bounds__(256) k_obs_update(Geo g, Dev d, int tag) {
    const int env = blockIdx.x;
    if (!d.obs_do[env] || d.fin[env] != tag) return;
    const int a = d.act[2 * env];
    if (a < 1) return;                                     // idle / stop: nothing moves
    const int n = d.n_remaining[env];                      // after the step
    const int *rank = d.rank_net + (size_t)env * g.max_nets;
    int ra = 0;                                            // old rank of a = remaining nets with a smaller id
    while (ra < n && rank[ra] < a) ra++;
    // clear: the routed net at its old block, every later net at its old block (new rank + 1)
    obs_mark_net(g, d, env, a, ra, 0.f);
    for (int k = ra; k < n; k++) obs_mark_net(g, d, env, rank[k], k + 1, 0.f);
    __syncthreads();
    // set: every later net at its new block
    for (int k = ra; k < n; k++) obs_mark_net(g, d, env, rank[k], k, 1.f);
    // order channel: n ids then a zero where the old last id was
    float *ch1 = d.obs + (size_t)env * g.obs_stride + g.cells;
    for (int k = ra + threadIdx.x; k <= n; k += blockDim.x) if (k < g.cells) ch1[k] = k < n ? (float)rank[k] : 0.f;
}

// Incremental reset, part 1 (before k_reset_env recomputes the rank list): remove the
// access-point ones and order ids of the nets that are still in the observation.
__global__ void __launch_bounds__(256) k_obs_reset_clear(Geo g, Dev d) {
    const int env = blockIdx.x;
    if (!d.obs_do[env] || d.obs_full[env]) return;
    const int n = d.n_remaining[env];
    const int *rank = d.rank_net + (size_t)env * g.max_nets;
    for (int k = 0; k < n; k++) obs_mark_net(g, d, env, rank[k], k, 0.f);
    float *ch1 = d.obs + (size_t)env * g.obs_stride + g.cells;
    for (int k = threadIdx.x; k < n; k += blockDim.x) if (k < g.cells) ch1[k] = 0.f;
}
// part 2 (after k_reset_cells / k_reset_env): channel 0 from the blockage bytes, every net
// at its initial rank, the order channel.  grid (chunks, N).
__global__ void __launch_bounds__(256) k_obs_reset_set(Geo g, Dev d) {
    const int env = blockIdx.y;
    if (!d.obs_do[env] || d.obs_full[env]) return;
    float *out = d.obs + (size_t)env * g.obs_stride;
    const uint8_t *ob = d.obst_obs + (size_t)env * g.cells_o;
    const int n4 = g.cells >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const uchar4 b = reinterpret_cast<const uchar4 *>(ob)[i];
        reinterpret_cast<float4 *>(out)[i] = make_float4(b.x ? 1.f : 0.f, b.y ? 1.f : 0.f, b.z ? 1.f : 0.f, b.w ? 1.f : 0.f);
    }
    if (blockIdx.x == 0) {
        for (int i = n4 * 4 + threadIdx.x; i < g.cells; i += blockDim.x) out[i] = ob[i] ? 1.f : 0.f;
        const int n = d.n_remaining[env];
        const int *rank = d.rank_net + (size_t)env * g.max_nets;
        for (int k = 0; k < n; k++) obs_mark_net(g, d, env, rank[k], k, 1.f);
        float *ch1 = out + g.cells;
        for (int k = threadIdx.x; k < n; k += blockDim.x) if (k < g.cells) ch1[k] = (float)rank[k];
    }
}

// ---------------------------------------------------------------------- metrics
// Congestion reduction over the occupancy field (the "reward kernel"): blocked
// cells (wire on a blockage), shorted cells (two nets on a cell, or a foreign wire
// on a pin), overflow (sum of usage beyond capacity 1).  Stands in for the
// simulator-side Request.reward_violation (net_ordering.proto:37).  Reads 4 bytes
// per cell (cellinfo); apnet only for occupied AP cells.
__global__ void __launch_bounds__(256) k_metrics(Geo g, Dev d, int pass) {
    const int env = blockIdx.y;
    if (pass != -2 && (d.act[2 * env] < 1 || !pass_selects(d, env, pass))) return;   // -2: every env (kernel bench)
    const size_t eoff = (size_t)env * g.cells_p;
    const uint4 *ci4 = reinterpret_cast<const uint4 *>(d.cellinfo + eoff);
    const int n4 = g.cells_p >> 2;
    unsigned blocked = 0, shorted = 0, overflow = 0;
    // four independent 16-byte loads in flight per thread and trip
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * stride;
            v[u] = i < n4 ? __ldg(ci4 + i) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t c[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            if (((c[0] | c[1] | c[2] | c[3]) & CI_USAGE_MASK) == 0u) continue;    // nothing routed here
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t us = (c[k] & CI_USAGE_MASK) >> CI_USAGE_SHIFT;
                if (us) {
                    blocked += (c[k] & CI_BLOCK) ? 1u : 0u;
                    bool sh = us >= 2u;
                    if (!sh && (c[k] & CI_ISAP))
                        sh = d.apnet[eoff + 4 * (size_t)(i0 + u * stride) + k] != (c[k] & CI_OWNER_MASK);
                    shorted += sh ? 1u : 0u;
                    overflow += us - 1u;
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        blocked += __shfl_xor_sync(0xFFFFFFFFu, blocked, off);
        shorted += __shfl_xor_sync(0xFFFFFFFFu, shorted, off);
        overflow += __shfl_xor_sync(0xFFFFFFFFu, overflow, off);
    }
    __shared__ unsigned sm[3][8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[0][w] = blocked; sm[1][w] = shorted; sm[2][w] = overflow; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned sum = 0;
        for (int k = 0; k < 8; k++) sum += sm[threadIdx.x][k];
        if (sum) atomicAdd(&d.msum[4 * env + threadIdx.x], sum);
    }
}

// Rank list (remaining net ids ascending), legal mask, n_remaining of one env.
__device__ void refresh_remaining(const Geo &g, const Dev &d, int env) {
    // single thread: at most max_nets iterations
    const int *ns = d.net_start + (size_t)env * (g.max_nets + 2);
    const uint8_t *routed = d.routed + (size_t)env * (g.max_nets + 1);
    uint8_t *legal = d.legal + (size_t)env * (g.max_nets + 1);
    int *rank = d.rank_net + (size_t)env * g.max_nets;
    int n = 0;
    legal[0] = 0;
    for (int k = 1; k <= g.max_nets; k++) {
        const bool rem = !routed[k] && ns[k + 1] > ns[k];
        legal[k] = rem;
        if (rem) rank[n++] = k;
    }
    d.n_remaining[env] = n;
}

// Per-step epilogue (baseline_utils.py:426-438): metric deltas, done flag, reward
// (train_PPO.py:101-102), remaining-net list for the order channel.
__global__ void k_finalize(Geo g, Dev d, int pass, int use_minc) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= g.N || !pass_selects(d, env, pass)) return;
    d.fin[env] = pass_tag(pass);
    const int raw = d.act[2 * env];
    if (raw < 1) {
        d.delta[3 * env] = 0; d.delta[3 * env + 1] = 0; d.delta[3 * env + 2] = 0;
        d.reward[env] = 0.0;
        if (raw == -1) {                                    // stopped (net_ordering.proto:48): no action is legal any more
            uint8_t *legal = d.legal + (size_t)env * (g.max_nets + 1);
            for (int k = 0; k <= g.max_nets; k++) legal[k] = 0;
        }
        return;
    }
    long long *cum = d.cum + 6 * (size_t)env;
    // congestion counts: the sums of this step's scan (k_metrics), or the counts the commits maintain
    const long long blocked = use_minc ? d.minc[4 * env] : (long long)d.msum[4 * env];
    const long long shorted = use_minc ? d.minc[4 * env + 1] : (long long)d.msum[4 * env + 1];
    const long long overflow = use_minc ? d.minc[4 * env + 2] : (long long)d.msum[4 * env + 2];
    d.msum[4 * env] = 0; d.msum[4 * env + 1] = 0; d.msum[4 * env + 2] = 0;
    const long long vio = blocked + shorted, wl = d.wlvia[2 * env], via = d.wlvia[2 * env + 1];
    const long long dv = vio - cum[0], dw = wl - cum[1], da = via - cum[2];
    d.delta[3 * env] = (int)dv; d.delta[3 * env + 1] = (int)dw; d.delta[3 * env + 2] = (int)da;
    {   // dynamic part of the routed net's feature vector: times routed, its last deltas
        float *nf = d.netfeat + ((size_t)env * (g.max_nets + 1) + raw) * XR_NF;
        nf[18] += 1.f; nf[19] = (float)dv; nf[20] = (float)dw; nf[21] = (float)da;
    }
    cum[0] = vio; cum[1] = wl; cum[2] = via; cum[3] = blocked; cum[4] = shorted; cum[5] = overflow;
    double r = -1.0;
    r *= (double)dv * 500 + (double)da * 4 + (double)dw * 0.5;
    d.reward[env] = r;
    refresh_remaining(g, d, env);
    const bool dn = d.n_remaining[env] == 0;
    d.done[env] = dn;
    long long *es = d.envstat + 8 * (size_t)env;
    es[0] += 1;
    if (dn) es[1] += 1;
    es[4] += dv; es[5] += dw; es[6] += da;
}

// ------------------------------------------------------------------------ reset
// Restore the environments flagged in obs_do to their loaded instance: clear
// occupancy, rebuild the obstacle-channel bytes (transposing router -> observation
// layout) and the per-env counters.  grid (chunks, N).
__global__ void __launch_bounds__(256) k_reset_cells(Geo g, Dev d) {
    const int env = blockIdx.y;
    if (!d.obs_do[env]) return;
    const size_t eoff = (size_t)env * g.cells_p;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.cells_p; i += stride) {
        const uint32_t ci = d.cellinfo[eoff + i] & CI_STATIC_MASK;
        d.cellinfo[eoff + i] = ci;
        // frontier cell word: no wire, stale distance, the static flags
        d.dist64[eoff + i] = FRW_STALE | ((ci & CI_BLOCK) ? FRW_BLK : 0ull) | ((ci & CI_ISAP) ? FRW_AP : 0ull);
    }
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < g.cells; o += stride) {
        const int z = o % g.Z, y = (o / g.Z) % g.Y, x = o / (g.Z * g.Y);
        const uint32_t ci = d.cellinfo[eoff + ((size_t)z * g.Y + y) * g.Xp + x];
        d.obst_obs[(size_t)env * g.cells_o + o] = (ci & CI_BLOCK) ? 1 : 0;
    }
}
__global__ void k_reset_env(Geo g, Dev d) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= g.N || !d.obs_do[env]) return;
    uint8_t *routed = d.routed + (size_t)env * (g.max_nets + 1);
    for (int k = 0; k <= g.max_nets; k++) routed[k] = 0;
    for (int k = 0; k <= g.max_nets; k++) {
        float *nf = d.netfeat + ((size_t)env * (g.max_nets + 1) + k) * XR_NF;
        nf[18] = 0.f; nf[19] = 0.f; nf[20] = 0.f; nf[21] = 0.f;
    }
    for (int k = 0; k < 6; k++) d.cum[6 * (size_t)env + k] = 0;
    d.wlvia[2 * env] = 0; d.wlvia[2 * env + 1] = 0;
    d.msum[4 * env] = 0; d.msum[4 * env + 1] = 0; d.msum[4 * env + 2] = 0;
    d.minc[4 * env] = 0; d.minc[4 * env + 1] = 0; d.minc[4 * env + 2] = 0;
    d.delta[3 * env] = 0; d.delta[3 * env + 1] = 0; d.delta[3 * env + 2] = 0;
    d.reward[env] = 0.0;
    d.phase[env] = 0; d.path_n[env] = 0; d.conn_n[env] = 0;
    d.fr_epoch[env] = 0;                                    // (k_reset_cells made every cell word stale)
    refresh_remaining(g, d, env);
    d.done[env] = d.n_remaining[env] == 0;
}

// Mark which environments a reset touches.  ids == nullptr: all.
__global__ void k_mark(Geo g, Dev d, int value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g.N) d.obs_do[i] = (uint8_t)value;
}
__global__ void k_mark_ids(Geo g, Dev d, const int *ids, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) d.obs_do[ids[i]] = 1;
}

// ------------------------------------------------------------------------ stats
// Sum the per-environment counters into the int64 vector that is all-reduced
// across GPUs (torch.distributed / NCCL, SUM).  One block.
__global__ void __launch_bounds__(256) k_stats(Geo g, Dev d) {
    long long acc[12];
#pragma unroll
    for (int k = 0; k < 12; k++) acc[k] = 0;
    for (int e = threadIdx.x; e < g.N; e += blockDim.x) {
        const long long *cum = d.cum + 6 * (size_t)e;
        const long long *es = d.envstat + 8 * (size_t)e;
        acc[0] += es[0]; acc[1] += es[1];
        acc[2] += es[4]; acc[3] += es[5]; acc[4] += es[6];     // lifetime sums of the step deltas
        acc[5] += cum[3]; acc[6] += cum[4]; acc[7] += cum[5];  // congestion snapshot of the running episodes
        acc[8] += -(es[4] * 1000 + es[6] * 8 + es[5]);         // 2 * lifetime reward (exact integer)
        acc[9] += es[2]; acc[10] += es[7]; acc[11] += es[3];
    }
    __shared__ long long sm[12][8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        long long v = acc[k];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
        if (lane == 0) sm[k][w] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        long long s = 0;
        if (threadIdx.x < 12) for (int k = 0; k < 8; k++) s += sm[threadIdx.x][k];
        d.stats[threadIdx.x] = s;
    }
}
