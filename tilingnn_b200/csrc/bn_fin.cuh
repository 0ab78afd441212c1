// Train-mode BatchNorm statistics without launches of their own (small graphs).
//
// Producers write ONE partial row per CTA (block_part_store: the warps' fp64 column sums are added in a fixed order in
// shared memory), and the CONSUMER of a BatchNorm'd tensor finishes the statistic in its own prologue
// (bn_finish_block: every CTA sums the partial rows in the same fixed order, so all CTAs get identical coefficients;
// CTA 0 publishes them).  At N ~ 600 (depth 20) a forward is ~75 dependent launches of a few microseconds each, and a
// k_bn_finish between every producer and consumer was a quarter of them.
// Reference semantics: nn.BatchNorm1d in train mode, biased variance over the rows of this call.
#pragma once

#include <cstdint>

#include "tgnn_internal.h"

namespace tgnn {

// Second half of block_part_store for callers that fill the scratch rows themselves ([warp][64] doubles, after a CTA barrier).
__device__ __forceinline__ void block_part_finish(double* __restrict__ part, const double* scratch, int nwarps) {
    __syncthreads();
    if (threadIdx.x < 64) {
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += scratch[w * 64 + threadIdx.x];
        part[(size_t)blockIdx.x * 64 + threadIdx.x] = t;
    }
}

// Adds the calling warps' per-lane column sums (lane = column, s1 = sum, s2 = sum of squares: a [64]-double row per warp)
// in warp order and stores ONE row per CTA: part[blockIdx.x][64].  scratch: shared memory, >= nwarps * 64 doubles, not in
// use by any warp of the CTA any more (the function synchronises the CTA before writing it).  Every thread of the CTA must call.
__device__ __forceinline__ void block_part_store(double* __restrict__ part, double s1, double s2, double* scratch, int nwarps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    scratch[warp * 64 + lane] = s1;
    scratch[warp * 64 + 32 + lane] = s2;
    block_part_finish(part, scratch, nwarps);
}

// Finishes one BatchNorm of C channels: cs (shared, [4][C] floats) receives the coefficients.  scratch: shared memory,
// >= nsl * 2 C doubles with nsl = max(1, min(8, blockDim.x / (2 C))).  Every thread of the CTA must call (two CTA barriers).
// The partial rows of a column are summed by nsl thread slices (every nsl-th row each, all loads in flight together) and
// the slices are added in slice order: the same association in every CTA.
__device__ __forceinline__ void bn_finish_block(const BnFin& f, int C, float* cs, double* scratch) {
    const int NC = 2 * C, T = blockDim.x;
    int nsl = T / NC;
    nsl = nsl < 1 ? 1 : (nsl > 8 ? 8 : nsl);
    double cnt = f.count;
    if (f.count_ptr) cnt = *f.count_ptr;
    for (int item = threadIdx.x; item < NC * nsl; item += T) {
        const int col = item % NC, slice = item / NC;
        const double* p = f.part + col;
        double s = 0.0;
        for (int r0 = slice; r0 < f.n_part; r0 += 16 * nsl) {          // 16 rows in flight, then their adds (row order)
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { const int r = r0 + k * nsl; v[k] = 0.0; if (r < f.n_part) v[k] = p[(size_t)r * NC]; }
#pragma unroll
            for (int k = 0; k < 16; ++k) s += v[k];
        }
        scratch[item] = s;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += T) {
        double t1 = 0.0, t2 = 0.0;
        for (int k = 0; k < nsl; ++k) { t1 += scratch[k * NC + c]; t2 += scratch[k * NC + C + c]; }
        const double mean = t1 / cnt;
        double var = t2 / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const double rstd = 1.0 / sqrt(var + BN_EPS);
        const float mh = (float)mean, ml = (float)(mean - (double)mh);
        const float sc = (float)((double)f.gamma[c] * rstd), be = f.beta[c];
        cs[c] = mh; cs[C + c] = ml; cs[2 * C + c] = sc; cs[3 * C + c] = be;
        if (blockIdx.x == 0 && blockIdx.y == 0 && f.coef_out) {
            f.coef_out[c] = mh; f.coef_out[C + c] = ml; f.coef_out[2 * C + c] = sc; f.coef_out[3 * C + c] = be;
        }
    }
    __syncthreads();
}

}  // namespace tgnn
