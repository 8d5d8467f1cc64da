// xr_frontier.h -- host interface of the frontier routing engine (xr_frontier.cu), used by xr_api.cu.
#pragma once
#include "xr_common.cuh"

#define FR_MAXAP 1024      // access points of one net cached on chip (larger nets take the full-grid sweeps)
#define FR_MAXPIN 264      // pins of one net (conn_cap + 1 = 257 is the ABI limit)
#ifndef FR_T
#define FR_T 1024          // largest block of the frontier kernel (register budget: 65536 / FR_T per thread)
#endif
#ifndef FR_MINB
#define FR_MINB 1
#endif
#ifndef FR_RAY
#define FR_RAY 8           // cells an expansion relaxes along the layer's preferred direction, at most (lanes per ray: 4, 8 or 16)
#endif

struct FrParams {
    int cap_s;             // open-list entries kept in shared memory (per list)
    int cap_e;             // expansion-list entries kept in shared memory
    int cap_g;             // open-list entries that may spill to global memory (per list)
    int cap_ge;            // expansion-list entries that may spill
    uint32_t delta;        // entries with f <= (smallest open f) + delta are expanded in the same round (cost units)
    int ray;               // 1..FR_RAY
    int dmax;              // the bucket may widen to dmax * delta while rounds are nearly empty (1 = fixed width)
    int park_min;          // open-list size from which entries far above the bucket are parked in a far list (0 = never)
    uint32_t band;         // the far list takes entries with f > (smallest open f) + band (cost units)
};

size_t xr_frontier_smem(const Geo &g, const FrParams &P);
size_t xr_frontier_spill_words(const FrParams &P);
cudaError_t xr_frontier_init(int smem_cap);
cudaError_t xr_frontier_launch(const Geo &g, const Dev &d, const int *env_list, int n_envs, const FrParams &P,
                               int threads, cudaStream_t st);
