// Scoring kernels of the TilinGNN forward pass for sm_100a (fp32 arithmetic, fp64 BatchNorm sums).
//
// Layout in HBM: every node tensor is row-major [rows][32] fp32, one 128-byte line per node, so a
// gather of a neighbour is exactly one coalesced line.  See DESIGN.md for the per-kernel byte model.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "conv_adj_body.cuh"
#include "gin_mlp.cuh"
#include "hsplit.cuh"
#include "layouts.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {

using namespace tfx;
using namespace ginx;
constexpr int WARPS = 8;        // warps per CTA in the warp-tile kernels
constexpr int TPB = WARPS * 32;

// 16-row x 32-col product on CUDA cores.  Thread (a = lane>>3, q = lane&7) owns rows a+4i (i<4) and
// columns 4q..4q+3.  xs: [16][XS] in shared memory, W: [KDIM][ldw] (k-major) readable with float4 loads.
template <int KDIM, bool W_GLOBAL>
__device__ __forceinline__ void tile16_fma(const float* __restrict__ xs, int xstride, const float* __restrict__ W,
                                           int ldw, int col0, int a, int q, float (&m)[4][4]) {
#pragma unroll
    for (int kk = 0; kk < KDIM / 4; ++kk) {
        float4 xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            xv[i] = *reinterpret_cast<const float4*>(xs + (a + 4 * i) * xstride + 4 * kk);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            const float4* wp = reinterpret_cast<const float4*>(W + (size_t)(4 * kk + k2) * ldw + col0) + q;
            float4 w = W_GLOBAL ? __ldg(wp) : *wp;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float x = k2 == 0 ? xv[i].x : k2 == 1 ? xv[i].y : k2 == 2 ? xv[i].z : xv[i].w;
                m[i][0] = fmaf(x, w.x, m[i][0]);
                m[i][1] = fmaf(x, w.y, m[i][1]);
                m[i][2] = fmaf(x, w.z, m[i][2]);
                m[i][3] = fmaf(x, w.w, m[i][3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Adjacency branch on fp32 rows with 3xTF32 (TGNN_CONV=chunk); the arithmetic lives in conv_adj_body.cuh because
// k_conv_h (conv_h.cu) falls through to it when an operand is outside the fp16 range.
// ------------------------------------------------------------------------------------------------
template <int WN, int NW>
__global__ void __launch_bounds__(NW * 32, WN == WN_BIG ? 1 : 2)
k_conv_adj(ConvArgs A) {
    extern __shared__ __align__(16) float smem[];
    conv_adj_body<WN, NW>(A, smem);
}

// ------------------------------------------------------------------------------------------------
// Collision branch: GINConv (sum over CSR neighbours + self) and its 32->32->64->32 sigmoid MLP,
// LeakyReLU, BatchNorm partial sums.  (graph_networks/layers/coll_conv.py:24-27; PyG GINConv.)
// The previous layer's BatchNorm is applied lazily on the gathered rows:
//   sum_j BN(x_j) = scale * sum_j ((x_j - mu_hi) - mu_lo) + (#terms) * beta
// Gather: four 8-lane groups each stream every 4th neighbour row with LDG.128 (4 rows per instruction,
// 16 in flight per warp).  MLP: three chained mma.sync 3xTF32 layers on 16-node chunks; layer k's C
// fragments are layer k+1's A fragments (KMAP_CHAIN), weights are frag tables in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int GIN_IDX_CAP = 1024;                                         // staged neighbour indices per 16-node chunk
constexpr int GIN_WARP_FLOATS = CH * XS + GIN_IDX_CAP;

__device__ __forceinline__ void add4(float4& a, const float4& b) { f4add(a, b); }     // two FADD2

// HMLP: layers 2 and 3 of the MLP on fp16-split operands (mma.sync m16n8k16: half the MMAs and half the weight-table
// reads of 3xTF32).  Their inputs are sigmoid outputs in (0, 1), always inside the fp16 range; layer 1 sees unbounded
// neighbour sums and stays on 3xTF32.  The host picks HMLP only when the layer's weights are inside the range too.
// weight buffer (global): [W1 | W2 | W3 | biases] (3xTF32 tables) then [W2h | W3h] (fp16 hi|lo tables).
template <bool HMLP>
__global__ void __launch_bounds__(TPB, 2)
k_gin(GinArgs A) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WF = HMLP ? GIN_WFLOATS_H : GIN_WFLOATS;
    gin_load_weights<HMLP>(smem, A.wfrag, threadIdx.x, TPB);
    __syncthreads();
    const GinW<HMLP> Wt(smem);
    float* xs = smem + WF + warp * GIN_WARP_FLOATS;
    int* sidx = reinterpret_cast<int*>(xs + CH * XS);
    const int a = lane >> 3, q = lane & 7;       // gather roles
    const int g = lane >> 2, t = lane & 3;       // mma roles
    const int gwarp = blockIdx.x * WARPS + warp, nwarp = gridDim.x * WARPS;
    const int n_chunks = (A.n_own + CH - 1) / CH;
    const float self_w = 1.0f + A.eps;
    double s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.0; s2[j] = 0.0; }

    for (int chunk = gwarp; chunk < n_chunks; chunk += nwarp) {
        const int node0 = chunk * CH;
        // ---- gather + sum: each 8-lane group owns nodes 4i+a of the chunk and keeps 8 neighbour rows in flight;
        // ---- the chunk's neighbour indices are staged in shared memory first (no dependent global load).  The rows
        // ---- are final values (h0 or the previous CollConv output written by k_combine): a plain fp32 sum --------
        int p = 0;
        if (lane <= CH) p = __ldg(A.col_ptr + min(node0 + lane, A.n_own));
        const int e_lo = __shfl_sync(0xffffffffu, p, 0), e_hi = __shfl_sync(0xffffffffu, p, CH);
        const bool staged = (e_hi - e_lo) <= GIN_IDX_CAP;
        if (staged) for (int i = lane; i < e_hi - e_lo; i += 32) sidx[i] = __ldg(A.col_src + e_lo + i);
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < CH / 4; ++i) {
            const int r = 4 * i + a, node = node0 + r;
            const int e0 = __shfl_sync(0xffffffffu, p, r), e1 = __shfl_sync(0xffffffffu, p, r + 1);
            const bool live = node < A.n_own;
            const int n_mine = live ? e1 - e0 : 0;
            int n_max = max(n_mine, __shfl_xor_sync(0xffffffffu, n_mine, 8));       // warp-uniform trip count
            n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, 16));
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float4 c = ld_row4(A.xin, node, q);
                sum.x = self_w * c.x; sum.y = self_w * c.y; sum.z = self_w * c.z; sum.w = self_w * c.w;
            }
            // indices first (clamped, branch-free: shared memory when staged, else global), then up to 8 predicated
            // row loads in flight, then the adds
            const int last = max(n_mine - 1, 0);
            for (int o = 0; o < n_max; o += 8) {
                int idx[8];
                if (staged) {
                    // the four lane groups walk a batch in rotated order (k + 2a): with equal degrees of 32 their lists
                    // are 32 words apart and the same position would be four words of ONE bank
                    const int* ip = sidx + (e0 - e_lo);
#pragma unroll
                    for (int k = 0; k < 8; ++k) idx[k] = ip[min(o + ((k + 2 * a) & 7), last)];
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) idx[k] = n_mine > 0 ? __ldg(A.col_src + e0 + min(o + ((k + 2 * a) & 7), last)) : 0;
                }
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (o + ((k + 2 * a) & 7) < n_mine) v[k] = ld_row4(A.xin, idx[k], q);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) add4(sum, v[k]);
            }
            *reinterpret_cast<float4*>(xs + r * XS + 4 * q) = sum;
        }
        __syncwarp();
        float a1[4][4];
        gin_load_a1(xs, lane, a1);
        gin_mlp_chunk<HMLP>(a1, Wt, node0, A.n_own, A.out, s1, s2, lane, A.mask);
        __syncwarp();
    }
    // fold the eight row groups (lanes with equal t) in a fixed order; lanes 0..3 publish channels 8t..8t+7
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
    }
    if (A.part) {                                 // one partial row per CTA: the warps' rows are added in warp order
        double* scratch = reinterpret_cast<double*>(smem);            // (the weight tables: no warp reads them any more)
        __syncthreads();
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                scratch[warp * 64 + 8 * t + j] = s1[j];
                scratch[warp * 64 + 32 + 8 * t + j] = s2[j];
            }
        }
        block_part_finish(A.part, scratch, WARPS);
    }
}

// ------------------------------------------------------------------------------------------------
// Collision branch on SMALL graphs ("S" variant, n_own <= GIN_S_MAX_NODES): the same arithmetic as k_gin, laid out for latency
// instead of throughput.  At N ~ 600 k_gin is 38 warps that each walk 4 rounds x 5 batches of dependent L2 loads and then
// the MLP (19 us per launch, a third of a depth-20 forward).  Here one CTA of four warps takes one 16-node chunk: every
// 8-lane group owns ONE destination row (a single round), all of the row's neighbour indices are fetched at once and
// 16 neighbour rows are in flight per lane; the MLP's weight tables arrive by cp.async while the gather runs; warp 0
// then runs the MLP.  One BatchNorm partial row per CTA.
// ------------------------------------------------------------------------------------------------
constexpr int GIN_S_THREADS = 128;
constexpr int GIN_S_MAX_NODES = 4096;           // 256 chunks (= 256 BatchNorm partial rows for the consumer's prologue); beyond that k_gin

template <bool HMLP>
__device__ __forceinline__ void gin_load_weights_async(float* smem, const float* __restrict__ wfrag, int tid, int nthreads) {
    auto cp16 = [](float* dst, const float* src) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    if (HMLP) {
        for (int i = tid; i < GIN_W1 / 4; i += nthreads) cp16(smem + 4 * i, wfrag + 4 * i);
        for (int i = tid; i < (GIN_W2H + GIN_W3H) / 4; i += nthreads) cp16(smem + GIN_W1 + 4 * i, wfrag + GIN_WFLOATS + 4 * i);
        for (int i = tid; i < 128 / 4; i += nthreads) cp16(smem + GIN_W1 + GIN_W2H + GIN_W3H + 4 * i, wfrag + GIN_W1 + GIN_W2 + GIN_W3 + 4 * i);
    } else {
        for (int i = tid; i < GIN_WFLOATS / 4; i += nthreads) cp16(smem + 4 * i, wfrag + 4 * i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <bool HMLP>
__global__ void __launch_bounds__(GIN_S_THREADS)
k_gin_s(GinArgs A) {
    extern __shared__ __align__(16) float smem[];
    constexpr int WF = HMLP ? GIN_WFLOATS_H : GIN_WFLOATS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    gin_load_weights_async<HMLP>(smem, A.wfrag, tid, GIN_S_THREADS);
    float* xs = smem + WF;                       // [CH][XS] neighbour sums of the chunk
    const int a = lane >> 3, q = lane & 7;
    const int node0 = blockIdx.x * CH;
    const int r = 4 * warp + a, node = node0 + r;
    const bool live = node < A.n_own;
    int e0 = 0, e1 = 0;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
        e0 = __ldg(A.col_ptr + node); e1 = __ldg(A.col_ptr + node + 1);
        const float4 c = ld_row4(A.xin, node, q);
        const float self_w = 1.0f + A.eps;
        sum.x = self_w * c.x; sum.y = self_w * c.y; sum.z = self_w * c.z; sum.w = self_w * c.w;
    }
    const int n_mine = e1 - e0;
    int n_max = max(n_mine, __shfl_xor_sync(0xffffffffu, n_mine, 8));       // warp-uniform trip count
    n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, 16));
    for (int o64 = 0; o64 < n_max; o64 += 64) {
        // lane q of the row's group fetches index positions q, q + 8, ... (<= 64 per pass), all in flight together
        int ix[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { ix[k] = 0; if (o64 + 8 * k + q < n_mine) ix[k] = __ldg(A.col_src + e0 + o64 + 8 * k + q); }
#pragma unroll
        for (int b = 0; b < 4; ++b) {            // 16 neighbour rows in flight per lane, then their adds (CSR order)
            if (o64 + 16 * b >= n_max) break;
            float4 v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int idx = __shfl_sync(0xffffffffu, ix[2 * b + (k >> 3)], (lane & 24) + (k & 7));
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o64 + 16 * b + k < n_mine) v[k] = ld_row4(A.xin, idx, q);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) add4(sum, v[k]);
        }
    }
    *reinterpret_cast<float4*>(xs + r * XS + 4 * q) = sum;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (warp != 0) return;
    const GinW<HMLP> Wt(smem);
    double s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
    float a1[4][4];
    gin_load_a1(xs, lane, a1);
    gin_mlp_chunk<HMLP>(a1, Wt, node0, A.n_own, A.out, s1, s2, lane, A.mask);
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
        }
    }
    if (A.part && g == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            A.part[(size_t)blockIdx.x * 64 + 8 * t + j] = s1[j];
            A.part[(size_t)blockIdx.x * 64 + 32 + 8 * t + j] = s2[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// combine: b1_new = BN_a(pre1) * BN_c(pre2) + middle[i-2]     (TilinGNN.py:64-69)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float bn_apply(float x, const float* __restrict__ coef, int c, int C) {
    return fmaf((x - coef[c]) - coef[C + c], coef[2 * C + c], coef[3 * C + c]);
}

template <bool FIN>
__global__ void __launch_bounds__(FIN ? 1024 : 256) k_combine(const float4* __restrict__ pre1, const float* __restrict__ coef1,
                          const float4* __restrict__ pre2, const float* __restrict__ coef2,
                          const float4* __restrict__ res, float4* __restrict__ out, uint4* __restrict__ xh,
                          int* __restrict__ flag, float4* __restrict__ g2out, int64_t n4, CombineFin fin, const uint8_t* __restrict__ mask) {
    __shared__ float c1[128], c2[128];
    bool bad = false;
    if (FIN) {
        // small graphs: every block finishes the two BatchNorms itself (same fixed order everywhere, so all blocks
        // get identical coefficients) instead of waiting for a separate k_bn_finish launch; block 0 publishes them.
        // The block's (blockDim / 128) thread slices each sum every nsl-th partial row of a column with all their loads in
        // flight at once (the prologue is one L2 latency, not one per 8 rows); the slices are added in a fixed order.
        __shared__ double ssum[8][128];
        const int nsl = blockDim.x >> 7;                                       // 8 at the launch width of 1024 threads
        double cnt = fin.count;
        if (fin.count_ptr && threadIdx.x < 64) cnt = *fin.count_ptr;           // (in flight with the partial rows)
        {
            const int col128 = threadIdx.x & 127, slice = threadIdx.x >> 7;
            const int w = col128 >> 6, col = col128 & 63;
            const double* p = (w == 0 ? fin.part[0] : fin.part[1]) + col;      // (no dynamic indexing of the parameter struct)
            const int nr = w == 0 ? fin.n_part[0] : fin.n_part[1];
            double s = 0.0;
            for (int r0 = slice; r0 < nr; r0 += 16 * nsl) {                    // 16 rows in flight, then their adds (row order)
                double v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) { const int r = r0 + k * nsl; v[k] = 0.0; if (r < nr) v[k] = p[(size_t)r * 64]; }
#pragma unroll
                for (int k = 0; k < 16; ++k) s += v[k];
            }
            ssum[slice][col128] = s;
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            const int w = threadIdx.x >> 5, c = threadIdx.x & 31;
            double t1 = 0.0, t2 = 0.0;
            for (int k = 0; k < nsl; ++k) { t1 += ssum[k][w * 64 + c]; t2 += ssum[k][w * 64 + 32 + c]; }
            const double mean = t1 / cnt;
            double var = t2 / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            const double rstd = 1.0 / sqrt(var + BN_EPS);
            const float mh = (float)mean, ml = (float)(mean - (double)mh);
            const float sc = (float)((double)(w == 0 ? fin.gamma[0] : fin.gamma[1])[c] * rstd), be = (w == 0 ? fin.beta[0] : fin.beta[1])[c];
            float* cs = w == 0 ? c1 : c2;
            cs[c] = mh; cs[32 + c] = ml; cs[64 + c] = sc; cs[96 + c] = be;
            if (blockIdx.x == 0) {
                float* cg = w == 0 ? fin.coef_out[0] : fin.coef_out[1];
                cg[c] = mh; cg[32 + c] = ml; cg[64 + c] = sc; cg[96 + c] = be;
            }
        }
    } else if (threadIdx.x < 128) { c1[threadIdx.x] = coef1[threadIdx.x]; c2[threadIdx.x] = coef2[threadIdx.x]; }
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i & 7) * 4;
        float4 p = __ldg(pre1 + i), g = __ldg(pre2 + i);
        float4 o, g2;
        g2.x = bn_apply(g.x, c2, c + 0, 32); g2.y = bn_apply(g.y, c2, c + 1, 32);
        g2.z = bn_apply(g.z, c2, c + 2, 32); g2.w = bn_apply(g.w, c2, c + 3, 32);
        if (g2out) g2out[i] = g2;                 // CollConv output: the next layer's k_gin gathers it as is
        o.x = bn_apply(p.x, c1, c + 0, 32) * g2.x;
        o.y = bn_apply(p.y, c1, c + 1, 32) * g2.y;
        o.z = bn_apply(p.z, c1, c + 2, 32) * g2.z;
        o.w = bn_apply(p.w, c1, c + 3, 32) * g2.w;
        if (res) { float4 r = __ldg(res + i); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
        if (!row_kept(mask, (int)(i >> 3))) {                   // masked node: zero rows, so every gather of them adds nothing
            o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g2out) g2out[i] = o;
        }
        out[i] = o;
        if (xh) xh[(i & ~(int64_t)7) + xh_pos((int)(i & 7))] = split_h4(o, bad);     // float4 q of a row -> uint4 xh_pos(q)
    }
    if (bad) *flag = 1;
}

// ------------------------------------------------------------------------------------------------
// init MLP (TilinGNN.py:31,54): two Linear -> LeakyReLU -> BN stages, recomputed from x per pass
// (MODE 0: statistics of stage 0, 1: statistics of stage 1, 2: write h0 and its fp16-split copy).
// One THREAD per node does both small matrix products out of registers (weights are shared-memory broadcasts:
// 8 LDS.128 feed 32 FFMA); the 32 nodes x 32 channels block of a warp is then transposed through shared memory
// so that lane = channel for the fp64 statistics (fixed order) and the coalesced row stores.
// ------------------------------------------------------------------------------------------------
constexpr int INIT_MAX_DX = 64;       // wider node features take k_init_wide (one warp per node)

template <int MODE>
__global__ void __launch_bounds__(TPB)
k_init(InitArgs A, const __grid_constant__ InitW1 W1) {
    __shared__ __align__(16) float w0t[INIT_MAX_DX * 32];      // [d][c]
    __shared__ float cf[2][128];
    __shared__ __align__(16) float tile[WARPS][32 * 33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < A.d_x * 32; i += TPB) w0t[i] = __ldg(A.w0 + (i & 31) * A.d_x + (i >> 5));
    const bool fin = MODE >= 1 && A.fin.part != nullptr;      // small graphs: this launch finishes the BatchNorm it needs fresh
    if (MODE >= 1) {
        if (threadIdx.x < 128 && !(fin && MODE == 1)) cf[0][threadIdx.x] = A.coef0[threadIdx.x];
    }
    if (MODE == 2 && threadIdx.x < 128 && !fin) cf[1][threadIdx.x] = A.coef1[threadIdx.x];
    if (fin) bn_finish_block(A.fin, 32, cf[MODE == 1 ? 0 : 1], reinterpret_cast<double*>(&tile[0][0]));
    __syncthreads();
    const float b0 = __ldg(A.b0 + lane), b1 = MODE >= 1 ? __ldg(A.b1 + lane) : 0.f;
    const int gwarp = blockIdx.x * WARPS + warp, nwarp = gridDim.x * WARPS;
    float* tl = tile[warp];
    double s1 = 0.0, s2 = 0.0;
    bool bad = false;
    for (int base = gwarp * 32; base < A.n_own; base += nwarp * 32) {
        const int node = base + lane;
        float v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = __shfl_sync(0xffffffffu, b0, c);
        if (node < A.n_own) {
            for (int d = 0; d < A.d_x; ++d) {
                const float xd = __ldg(A.x + (size_t)node * A.d_x + d);
                const float4* w = reinterpret_cast<const float4*>(w0t + d * 32);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 wv = w[c4];
                    v[4 * c4 + 0] = fmaf(xd, wv.x, v[4 * c4 + 0]); v[4 * c4 + 1] = fmaf(xd, wv.y, v[4 * c4 + 1]);
                    v[4 * c4 + 2] = fmaf(xd, wv.z, v[4 * c4 + 2]); v[4 * c4 + 3] = fmaf(xd, wv.w, v[4 * c4 + 3]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = leaky(v[c]);
        if (MODE >= 1) {
            float2 o[16];                           // channel pairs: packed FFMA2 (the same bits as 32 scalar FFMAs, half the issue slots)
#pragma unroll
            for (int c = 0; c < 16; ++c) o[c] = make_float2(__shfl_sync(0xffffffffu, b1, 2 * c), __shfl_sync(0xffffffffu, b1, 2 * c + 1));
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float y = fmaf((v[k] - cf[0][k]) - cf[0][32 + k], cf[0][64 + k], cf[0][96 + k]);
                const float2 yy = make_float2(y, y);
#pragma unroll
                for (int c = 0; c < 16; ++c)        // weights = kernel parameter (constant cache), read as pairs
                    o[c] = f2fma(yy, make_float2(W1.w[k * 32 + 2 * c], W1.w[k * 32 + 2 * c + 1]), o[c]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) { v[2 * c] = leaky(o[c].x); v[2 * c + 1] = leaky(o[c].y); }
        }
        // transpose: thread (node) major -> lane = channel
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) tl[lane * 33 + c] = v[c];
        __syncwarp();
        const int n_here = min(32, A.n_own - base);
        const unsigned keepbits = __ballot_sync(0xffffffffu, node >= A.n_own || row_kept(A.mask, node));    // bit j: node base + j is kept
        for (int j = 0; j < n_here; ++j) {
            const bool kept = (keepbits >> j) & 1u;
            const float val = kept ? tl[j * 33 + lane] : 0.f;
            if (MODE == 2) {
                const float o2 = kept ? fmaf((val - cf[1][lane]) - cf[1][32 + lane], cf[1][64 + lane], cf[1][96 + lane]) : 0.f;
                A.out[(size_t)(base + j) * F + lane] = o2;
                if (A.xh) {
                    // lane = channel -> word w of the split row: q = w>>2, {hi(4q,4q+1), hi(4q+2,4q+3), lo(..), lo(..)}[w&3]
                    __half hi, lo;
                    split_h(o2, hi, lo);
                    const uint32_t h16 = __half_as_ushort(hi), l16 = __half_as_ushort(lo);
                    const uint32_t wh = h16 | (__shfl_down_sync(0xffffffffu, h16, 1) << 16);
                    const uint32_t wl = l16 | (__shfl_down_sync(0xffffffffu, l16, 1) << 16);
                    const int src = (lane & ~3) + 2 * (lane & 1);
                    const uint32_t a = __shfl_sync(0xffffffffu, wh, src), b = __shfl_sync(0xffffffffu, wl, src);
                    A.xh[(size_t)(base + j) * F + 4 * xh_pos(lane >> 2) + (lane & 3)] = (lane & 2) ? b : a;
                    bad |= !(fabsf(o2) <= TG_H_LIMIT);
                }
            } else {
                s1 += (double)val;
                s2 += (double)val * (double)val;
            }
        }
    }
    if (MODE == 2 && bad) *A.flag = 1;
    if (MODE < 2 && A.part) block_part_store(A.part, s1, s2, reinterpret_cast<double*>(&tile[0][0]), WARPS);     // one partial row per CTA
}

// the same for d_x > 64: lane = channel, one warp per node
template <int MODE>
__global__ void __launch_bounds__(TPB)
k_init_wide(InitArgs A) {
    __shared__ __align__(16) float w1t[32 * 33];
    __shared__ float cfs[128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool fin = MODE >= 1 && A.fin.part != nullptr;      // small graphs: this launch finishes the BatchNorm it needs fresh
    if (fin) bn_finish_block(A.fin, 32, cfs, reinterpret_cast<double*>(w1t));
    const float* coef0 = (fin && MODE == 1) ? cfs : A.coef0;
    const float* coef1 = (fin && MODE == 2) ? cfs : A.coef1;
    if (MODE >= 1) {
        for (int i = threadIdx.x; i < 32 * 32; i += TPB) w1t[(i >> 5) * 33 + (i & 31)] = __ldg(A.w1t + i);
        __syncthreads();
    }
    const float b0 = __ldg(A.b0 + lane);
    const float b1 = MODE >= 1 ? __ldg(A.b1 + lane) : 0.f;
    const int gwarp = blockIdx.x * WARPS + warp, nwarp = gridDim.x * WARPS;
    double s1 = 0.0, s2 = 0.0;
    for (int node = gwarp; node < A.n_own; node += nwarp) {
        float v = b0;
        for (int d = 0; d < A.d_x; ++d) v = fmaf(__ldg(A.x + (size_t)node * A.d_x + d), __ldg(A.w0 + lane * A.d_x + d), v);
        v = leaky(v);
        const bool kept = row_kept(A.mask, node);
        if (!kept) v = 0.f;
        if (MODE >= 1) {
            float y = bn_apply(v, coef0, lane, 32);
            float o = b1;
#pragma unroll
            for (int k = 0; k < 32; ++k) o = fmaf(__shfl_sync(0xffffffffu, y, k), w1t[k * 33 + lane], o);
            v = kept ? leaky(o) : 0.f;
            if (MODE == 2) {
                const float o2 = kept ? bn_apply(v, coef1, lane, 32) : 0.f;
                A.out[(size_t)node * F + lane] = o2;
                if (A.xh) {
                    __half hi, lo;
                    split_h(o2, hi, lo);
                    const uint32_t h16 = __half_as_ushort(hi), l16 = __half_as_ushort(lo);
                    const uint32_t wh = h16 | (__shfl_down_sync(0xffffffffu, h16, 1) << 16);
                    const uint32_t wl = l16 | (__shfl_down_sync(0xffffffffu, l16, 1) << 16);
                    const int src = (lane & ~3) + 2 * (lane & 1);
                    const uint32_t a = __shfl_sync(0xffffffffu, wh, src), b = __shfl_sync(0xffffffffu, wl, src);
                    A.xh[(size_t)node * F + 4 * xh_pos(lane >> 2) + (lane & 3)] = (lane & 2) ? b : a;
                    if (!(fabsf(o2) <= TG_H_LIMIT)) *A.flag = 1;
                }
                continue;
            }
        }
        s1 += (double)v;
        s2 += (double)v * (double)v;
    }
    if (MODE < 2 && A.part) block_part_store(A.part, s1, s2, reinterpret_cast<double*>(w1t), WARPS);            // one partial row per CTA
}

// ------------------------------------------------------------------------------------------------
// Dense stage of the final MLP (TilinGNN.py:45-46,74-76): out = LeakyReLU(BN_in(A) @ Wt + b) and
// per-column partial sums.  A is either a plain [n][K] matrix or the virtual concat of K/32 slabs.
// CTA tile 128 rows x BN columns, k-chunks of 32, thread tile 8 x (BN/16).
// ------------------------------------------------------------------------------------------------
constexpr int DM = 128, DK = 32;

template <int BN>
__global__ void __launch_bounds__(256)
k_dense(DenseArgs A) {
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[DK][DM + 4];
    __shared__ __align__(16) float Bs[DK][BN];
    __shared__ double red[2][BN];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * DM, col0 = blockIdx.y * BN;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < A.K; k0 += DK) {
        const float* abase; int lda, koff;
        if (A.virtual_concat) { abase = A.slabs[k0 / DK]; lda = F; koff = 0; }
        else { abase = A.a; lda = A.K; koff = k0; }
        // A tile: 128 rows x 32 k; thread loads float4 (row = tid/8 + 32*j, k4 = tid%8)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = (tid >> 3) + 32 * j, k4 = (tid & 7) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < A.n) v = __ldg(reinterpret_cast<const float4*>(abase + (size_t)(row0 + r) * lda + koff + k4));
            if (A.in_coef) {
                const float* cf = A.in_coef; const int C = A.K, c = k0 + k4;
                v.x = bn_apply(v.x, cf, c + 0, C); v.y = bn_apply(v.y, cf, c + 1, C);
                v.z = bn_apply(v.z, cf, c + 2, C); v.w = bn_apply(v.w, cf, c + 3, C);
            }
            As[k4 + 0][r] = v.x; As[k4 + 1][r] = v.y; As[k4 + 2][r] = v.z; As[k4 + 3][r] = v.w;
        }
        // B tile: 32 k x BN cols
        for (int i = tid; i < DK * BN / 4; i += 256) {
            const int k = i / (BN / 4), c4 = (i % (BN / 4)) * 4;
            *reinterpret_cast<float4*>(&Bs[k][c4]) =
                __ldg(reinterpret_cast<const float4*>(A.wt + (size_t)(k0 + k) * A.n_out + col0 + c4));
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < DK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue
    if (tid < BN) { red[0][tid] = 0.0; red[1][tid] = 0.0; }
    __syncthreads();
    double cs1[TN], cs2[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) { cs1[j] = 0.0; cs2[j] = 0.0; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = row0 + ty * 8 + i;
        if (row < A.n) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int col = col0 + tx * TN + j;
                float v = leaky(acc[i][j] + __ldg(A.bias + col));
                if (!row_kept(A.mask, row)) v = 0.f;
                A.out[(size_t)row * A.n_out + col] = v;
                cs1[j] += (double)v; cs2[j] += (double)v * (double)v;
            }
        }
    }
    // deterministic column reduction over the 16 row groups: ty-ordered accumulation in shared memory
    for (int t = 0; t < 16; ++t) {
        if (ty == t) {
#pragma unroll
            for (int j = 0; j < TN; ++j) { red[0][tx * TN + j] += cs1[j]; red[1][tx * TN + j] += cs2[j]; }
        }
        __syncthreads();
    }
    if (A.part && tid < BN) {
        double* p = A.part + (size_t)blockIdx.x * 2 * A.n_out;
        p[col0 + tid] = red[0][tid];
        p[A.n_out + col0 + tid] = red[1][tid];
    }
}

// final Linear(32 -> 1) + Sigmoid on BN(a3)   (TilinGNN.py:47)
__global__ void k_score(const float* __restrict__ a3, const float* __restrict__ coef, const float* __restrict__ w,
                        float b, float* __restrict__ out, int64_t n, const uint8_t* __restrict__ mask, BnFin fin) {
    const int lane = threadIdx.x & 31;
    __shared__ float cfs[128];
    __shared__ double fin_scratch[8 * 64];
    if (fin.part) { bn_finish_block(fin, 32, cfs, fin_scratch); coef = cfs; }     // small graphs: the last BatchNorm is finished here
    const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float wl = __ldg(w + lane);
    // lane = channel: its four BatchNorm coefficients live in registers; a warp takes 8 consecutive rows per step (8 row loads
    // in flight, 8 interleaved shuffle trees) -- one row at a time was a chain of dependent L2 latencies (0.089 ms for 1M rows)
    const float mh = coef[lane], ml = coef[32 + lane], sc = coef[64 + lane], be = coef[96 + lane];
    for (int64_t base = gwarp * 8; base < n; base += nwarp * 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = 0.f; if (base + k < n) v[k] = __ldg(a3 + (base + k) * F + lane); }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaf((v[k] - mh) - ml, sc, be) * wl;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        float mine = v[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) if (lane == k) mine = v[k];
        const int64_t node = base + lane;
        if (lane < 8 && node < n) out[node] = row_kept(mask, (int)node) ? sigmoidf_acc(mine + b) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics
// ------------------------------------------------------------------------------------------------
__global__ void k_bn_reduce(const double* __restrict__ part, int n_part, int C2, double* __restrict__ sums) {
    __shared__ double sh[256];
    const int j = blockIdx.x;                        // one block per column of [sum | sumsq]
    double s = 0.0;
    for (int p = threadIdx.x; p < n_part; p += 256) s += part[(size_t)p * C2 + j];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[j] = sh[0];
}

__global__ void k_bn_coef(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ coef, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = sums[c] / count;
    double var = sums[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + BN_EPS);
    const float mh = (float)mean;
    coef[c] = mh;
    coef[C + c] = (float)(mean - (double)mh);
    coef[2 * C + c] = (float)((double)gamma[c] * rstd);
    coef[3 * C + c] = beta[c];
}

// Single-GPU train mode: reduce the partials of one or two BatchNorms AND turn them into coefficients in one launch.
// One block per column (fixed reduction order, as k_bn_reduce); the last block to finish (ticket) computes the
// coefficients from the published sums, so a layer's two BatchNorms cost one launch instead of four.
__global__ void k_bn_finish(BnFinishArgs A) {
    __shared__ double sh[256];
    __shared__ bool last;
    const int C2 = 2 * A.C;
    const int which = blockIdx.x / C2, j = blockIdx.x - which * C2;
    const double* part = A.part[which];
    const int n_part = A.n_part[which];
    double s = 0.0;
    for (int p = threadIdx.x; p < n_part; p += 256) s += part[(size_t)p * C2 + j];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        A.sums[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int nb = gridDim.x / C2;
    for (int i = threadIdx.x; i < nb * A.C; i += 256) {
        const int w = i / A.C, c = i - w * A.C;
        const volatile double* sums = A.sums + (size_t)w * C2;
        const double cnt = A.count_ptr ? *A.count_ptr : A.count;
        const double mean = sums[c] / cnt;
        double var = sums[A.C + c] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const double rstd = 1.0 / sqrt(var + BN_EPS);
        const float mh = (float)mean;
        float* coef = A.coef[w];
        coef[c] = mh;
        coef[A.C + c] = (float)(mean - (double)mh);
        coef[2 * A.C + c] = (float)((double)A.gamma[w][c] * rstd);
        coef[3 * A.C + c] = A.beta[w][c];
    }
    if (threadIdx.x == 0) *A.ticket = 0u;
}

// ---- peer-memory variants (sharded mode) ---------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Spin until flags[q] >= epoch for every peer q.  A peer's host may lag (first-call module load, garbage collection,
// CPU work between forwards), so the bound is generous (~30 s); past it the exchange cannot be trusted any more
// (a fast rank running on would break the "at most one exchange ahead" double-buffer invariant), so the kernel stores
// an error code in the handle's mapped host word and ABORTS: the stream's next operation fails and tgnn_forward /
// tgnn_check_error report it -- no NaN results, no later epochs.
__device__ __forceinline__ void wait_peers(const unsigned* flags, int world, int rank, unsigned epoch, int* err, unsigned peers = 0xffffffffu) {
    for (int q = 0; q < world; ++q) {
        if (q == rank || !((peers >> q) & 1u)) continue;
        long long t0 = clock64();
        while ((int)(ld_acquire_sys(flags + q) - epoch) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > 60000000000ll) {
                if (err) { *reinterpret_cast<volatile int*>(err) = TGNN_DEVERR_PEER; __threadfence_system(); }
                __trap();
            }
        }
    }
}

__global__ void k_bn_finish_x(BnFinishArgs A, PeerPtrs P, unsigned epoch) {
    __shared__ double sh[256];
    __shared__ bool last;
    const int C2 = 2 * A.C;
    const int which = blockIdx.x / C2, j = blockIdx.x - which * C2;
    const double* part = A.part[which];
    const int n_part = A.n_part[which];
    double s = 0.0;
    for (int p = threadIdx.x; p < n_part; p += 256) s += part[(size_t)p * C2 + j];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        A.sums[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const int n = gridDim.x, par = epoch & 1u;                 // n = n_bn * 2C local sums
    const size_t slot = (size_t)(par * PX_MAX_WORLD + P.rank) * PX_BN_SLOT;
    const volatile double* loc = A.sums;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double v = loc[i];
        for (int q = 0; q < P.world; ++q)                      // own copy too: the sum below reads one buffer only
            reinterpret_cast<double*>(P.base[q] + PX_FLAG_BYTES)[slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P.world && threadIdx.x != P.rank)
        st_release_sys(reinterpret_cast<unsigned*>(P.base[threadIdx.x] + 256) + par * PX_MAX_WORLD + P.rank, epoch);
    if (threadIdx.x == 0) wait_peers(reinterpret_cast<const unsigned*>(P.base[P.rank] + 256) + par * PX_MAX_WORLD, P.world, P.rank, epoch, P.err);
    __syncthreads();
    __threadfence_system();
    const double* mine = reinterpret_cast<const double*>(P.base[P.rank] + PX_FLAG_BYTES) + (size_t)par * PX_MAX_WORLD * PX_BN_SLOT;
    for (int i = threadIdx.x; i < n; i += 256) {
        double t = 0.0;
        for (int q = 0; q < P.world; ++q) t += __ldcg(mine + (size_t)q * PX_BN_SLOT + i);      // fixed rank order
        A.sums[i] = t;
    }
    __syncthreads();
    __threadfence();
    const int nb = n / C2;
    for (int i = threadIdx.x; i < nb * A.C; i += 256) {
        const int w = i / A.C, c = i - w * A.C;
        const volatile double* sums = A.sums + (size_t)w * C2;
        const double cnt = A.count_ptr ? *A.count_ptr : A.count;
        const double mean = sums[c] / cnt;
        double var = sums[A.C + c] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const double rstd = 1.0 / sqrt(var + BN_EPS);
        const float mh = (float)mean;
        float* coef = w == 0 ? A.coef[0] : A.coef[1];
        coef[c] = mh;
        coef[A.C + c] = (float)(mean - (double)mh);
        coef[2 * A.C + c] = (float)((double)(w == 0 ? A.gamma[0] : A.gamma[1])[c] * rstd);
        coef[3 * A.C + c] = (w == 0 ? A.beta[0] : A.beta[1])[c];
    }
    if (threadIdx.x == 0) *A.ticket = 0u;
}

// smask (optional): bit q of smask[r] = peer q reads boundary row r -- the row then travels only to those peers (node-range
// shards of a spatially ordered graph have 2 neighbours, not world - 1)
__global__ void k_halo_push(const float4* __restrict__ a, const float4* __restrict__ b, const int* __restrict__ rows, int n_send,
                            int64_t halo_slot, PeerPtrs P, unsigned epoch, unsigned* __restrict__ ticket, const uint8_t* __restrict__ smask) {
    const int par = epoch & 1u;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)n_send * 16) {
        const int r = (int)(i >> 4), c = (int)(i & 15);
        const int row = rows[r];
        const float4 v = c < 8 ? __ldg(a + (size_t)row * 8 + c) : (b ? __ldg(b + (size_t)row * 8 + (c - 8)) : make_float4(0.f, 0.f, 0.f, 0.f));
        const size_t off = ((size_t)par * P.world * halo_slot + (size_t)P.rank * halo_slot) * 16 + (size_t)i;     // float4 units
        const unsigned to = smask ? (unsigned)__ldg(smask + r) : 0xffffffffu;
        for (int q = 0; q < P.world; ++q)
            if (q != P.rank && ((to >> q) & 1u)) reinterpret_cast<float4*>(P.base[q] + PX_HALO_OFF)[off] = v;
    }
    // last block: everything this rank wrote is visible system-wide before the flags go up
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x < P.world && threadIdx.x != P.rank)
        st_release_sys(reinterpret_cast<unsigned*>(P.base[threadIdx.x]) + par * PX_MAX_WORLD + P.rank, epoch);
    if (threadIdx.x == 0) *ticket = 0u;
}

// used (optional): used[r] != 0 = mirrored row r is read by a local edge; need_from: peers that own such rows -- only their
// flags are waited for and only those rows are unpacked (rows nobody sent stay stale and are never read)
__global__ void k_halo_unpack_x(PeerPtrs P, unsigned epoch, int64_t halo_slot, int64_t n_own,
                                float4* __restrict__ a, float4* __restrict__ b, uint4* __restrict__ xh, int* __restrict__ flag,
                                const uint8_t* __restrict__ used, unsigned need_from) {
    const int par = epoch & 1u;
    if (threadIdx.x == 0) wait_peers(reinterpret_cast<const unsigned*>(P.base[P.rank]) + par * PX_MAX_WORLD, P.world, P.rank, epoch, P.err, need_from);
    __syncthreads();
    __threadfence_system();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)P.world * halo_slot * 16;
    if (i >= total) return;
    const int64_t r = i >> 4; const int c = (int)(i & 15);
    if (r / halo_slot == P.rank) return;               // own slot: rows are read in place
    if (used && !__ldg(used + r)) return;
    const float4* recv = reinterpret_cast<const float4*>(P.base[P.rank] + PX_HALO_OFF) + (size_t)par * total;
    float4 v = __ldcg(recv + i);                       // written by a peer over NVLink: L2 is the coherence point
    if (c < 8) {
        a[(size_t)(n_own + r) * 8 + c] = v;
        if (xh) {
            bool bad = false;
            xh[(size_t)(n_own + r) * 8 + xh_pos(c)] = split_h4(v, bad);
            if (bad) *flag = 1;
        }
    } else if (b) b[(size_t)(n_own + r) * 8 + (c - 8)] = v;
}

__global__ void k_bn_coef_eval(const float* __restrict__ rmean, const float* __restrict__ rvar,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* __restrict__ coef, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    coef[c] = rmean[c];
    coef[C + c] = 0.f;
    coef[2 * C + c] = (float)((double)gamma[c] / sqrt((double)rvar[c] + BN_EPS));
    coef[3 * C + c] = beta[c];
}

__global__ void k_transpose(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    int r = i / cols, c = i - r * cols;
    out[(size_t)c * rows + r] = in[i];
}

// frag table of a k-major [K][N] fp32 matrix (root weights, GIN MLP weights)
__global__ void k_frag_pack(const float* __restrict__ w, int K, int N, int kmap, int nmap, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    frag_store(out, i / N, i % N, N, kmap, nmap, (double)w[i]);
}

// fp16 hi|lo fragment table (natural k order) of a k-major [K][N] matrix; raises *flag outside the fp16 range
__global__ void k_frag_pack_h16(const float* __restrict__ w, int K, int N, int nmap, __half* __restrict__ out, int* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * N) return;
    const float v = w[i];
    __half hi, lo;
    split_h(v, hi, lo);
    out[hfrag_nat_half_index(i / N, i % N, N, nmap, 0)] = hi;
    out[hfrag_nat_half_index(i / N, i % N, N, nmap, 1)] = lo;
    if (!(fabsf(v) <= TG_H_LIMIT)) *flag = 1;
}

// ------------------------------------------------------------------------------------------------
// halo pack / unpack: rows of two [*, 32] tensors <-> [slot rows][64] exchange buffer
// ------------------------------------------------------------------------------------------------
__global__ void k_halo_pack(const float4* __restrict__ a, const float4* __restrict__ b, const int* __restrict__ rows,
                            int n_send, float4* __restrict__ sendbuf) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_send * 16) return;
    int r = (int)(i >> 4), c = (int)(i & 15);
    int row = rows[r];
    sendbuf[i] = c < 8 ? __ldg(a + (size_t)row * 8 + c) : (b ? __ldg(b + (size_t)row * 8 + (c - 8)) : make_float4(0.f, 0.f, 0.f, 0.f));
}

__global__ void k_halo_unpack(const float4* __restrict__ recv, int world, int rank, int64_t halo_slot, int64_t n_own,
                              float4* __restrict__ a, float4* __restrict__ b, uint4* __restrict__ xh, int* __restrict__ flag) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)world * halo_slot * 16;
    if (i >= total) return;
    int64_t r = i >> 4; int c = (int)(i & 15);
    if (r / halo_slot == rank) return;               // own slot: rows are read in place
    float4 v = __ldg(recv + i);
    if (c < 8) {
        a[(size_t)(n_own + r) * 8 + c] = v;
        if (xh) {                                    // mirrored rows feed k_conv_h too
            bool bad = false;
            xh[(size_t)(n_own + r) * 8 + xh_pos(c)] = split_h4(v, bad);
            if (bad) *flag = 1;
        }
    } else if (b) b[(size_t)(n_own + r) * 8 + (c - 8)] = v;
}

// ---- node mask (tgnn_set_node_mask): what changes when only the nodes with keep != 0 are part of the graph ---------
// per 64/128-row warp tile: kept in-degree of every destination (mean aggregation divides by it) and the number of
// adjacency edges with both endpoints kept
__global__ void __launch_bounds__(128)
k_mask_adj(const int* __restrict__ cptr, const int* __restrict__ csrc, const uint8_t* __restrict__ cdst, int wn, int n_own,
           const uint8_t* __restrict__ keep, float* __restrict__ inv_deg, int* __restrict__ counters) {
    __shared__ int deg[WN_BIG];
    const int tile = blockIdx.x;
    for (int i = threadIdx.x; i < wn; i += 128) deg[i] = 0;
    __syncthreads();
    const int64_t s0 = (int64_t)cptr[tile] * CH, s1 = (int64_t)cptr[tile + 1] * CH;
    int both = 0;
    for (int64_t s = s0 + threadIdx.x; s < s1; s += 128) {
        const int src = csrc[s];
        if (src >= 0 && keep[src]) {
            const int d = cdst[s];
            atomicAdd(&deg[d], 1);
            both += keep[tile * wn + d] ? 1 : 0;
        }
    }
    if (both) atomicAdd(counters + 1, both);
    __syncthreads();
    for (int i = threadIdx.x; i < wn; i += 128) {
        const int node = tile * wn + i;
        if (node < n_own) inv_deg[node] = 1.0f / (float)(deg[i] > 1 ? deg[i] : 1);
    }
}
// kept nodes and collision edges with both endpoints kept
__global__ void k_mask_count(const int* __restrict__ col_ptr, const int* __restrict__ col_src, int n_own,
                             const uint8_t* __restrict__ keep, int* __restrict__ counters) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    int kept = 0, both = 0;
    if (node < n_own && keep[node]) {
        kept = 1;
        for (int e = col_ptr[node]; e < col_ptr[node + 1]; ++e) both += keep[col_src[e]] ? 1 : 0;
    }
    kept = __reduce_add_sync(0xffffffffu, kept);
    both = __reduce_add_sync(0xffffffffu, both);
    if ((threadIdx.x & 31) == 0) {
        if (kept) atomicAdd(counters, kept);
        if (both) atomicAdd(counters + 2, both);
    }
}
__global__ void k_mask_finish(const int* __restrict__ counters, double* __restrict__ count) { *count = (double)counters[0]; }

int persistent_blocks(int work_items_per_block_unit, int sm_count, int blocks_per_sm) {
    int want = work_items_per_block_unit;
    int cap = sm_count * blocks_per_sm;
    return want < 1 ? 1 : (want < cap ? want : cap);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static size_t gin_smem(bool hmlp) { return (size_t)((hmlp ? GIN_WFLOATS_H : GIN_WFLOATS) + WARPS * GIN_WARP_FLOATS) * sizeof(float); }

ConvGeom conv_geom(int n_tiles, int wn, int sm_count, int64_t n_chunks) {
    ConvGeom g{};
    g.cluster = 1;
    if (wn == WN_BIG) {
        g.warps = 12; g.split = false;
        g.blocks = persistent_blocks((n_tiles + 11) / 12, sm_count, 1);
    } else {
        g.warps = WARPS;
        // few tiles per warp: one CTA per tile, its chunks split over the CTA's warps (a warp that owns a whole tile walks
        // ~150 chunks of ~2000 cycles each at deg 32: with fewer than ~3 tiles per warp the persistent geometry is one long
        // latency chain with idle SMs next to it)
        // Measured (B200, forward time, split vs persistent): 50k x deg 32 (782 tiles) -12 %, 100k x deg 32 (1563 tiles) -4 %,
        // 100k x deg 8 +4 %, 300k x deg 32 +13 %, 1M +10 % -> up to ~11 tiles per SM when the tiles are long (>= 100 chunks).
        const bool long_tiles = n_tiles > 0 && n_chunks >= (int64_t)100 * n_tiles;
        const int split_max = getenv("TGNN_CONV_SPLIT_MAX") ? atoi(getenv("TGNN_CONV_SPLIT_MAX")) : (long_tiles ? 11 : 2) * sm_count;
        g.split = n_tiles <= split_max;
        g.blocks = g.split ? (n_tiles < 1 ? 1 : n_tiles) : persistent_blocks((n_tiles + WARPS - 1) / WARPS, sm_count, 2);
        // fewer tiles than SMs: a tile's chunks are shared by a CLUSTER of 2 / 4 / 8 CTAs (one SM works through a chunk every
        // ~135 cycles however many warps it has, so the only way to finish a tile sooner is more SMs); partial tiles are
        // summed through distributed shared memory
        const bool cl_off = getenv("TGNN_CONV_CLUSTER") && std::string(getenv("TGNN_CONV_CLUSTER")) == "0";
        if (g.split && !cl_off) {
            const int room = sm_count / (n_tiles < 1 ? 1 : n_tiles);
            g.cluster = room >= 8 ? 8 : (room >= 4 ? 4 : (room >= 2 ? 2 : 1));
            g.blocks *= g.cluster;
        }
    }
    return g;
}
static int gin_blocks(int n_own, int sm_count) {
    int chunks = (n_own + CH - 1) / CH;
    return persistent_blocks((chunks + WARPS - 1) / WARPS, sm_count, 2);
}
static int init_blocks(int n_own, int sm_count) { return persistent_blocks((n_own + WARPS * 32 - 1) / (WARPS * 32), sm_count, 4); }

bool gin_small(int n_own) { return n_own <= GIN_S_MAX_NODES && !(getenv("TGNN_GIN_S") && std::string(getenv("TGNN_GIN_S")) == "0"); }
int gin_num_parts(int n_own, int sm_count) { return gin_small(n_own) ? (n_own + CH - 1) / CH : gin_blocks(n_own, sm_count); }   // one partial row per CTA
int init_num_parts(int n_own, int sm_count) { return init_blocks(n_own, sm_count); }
int dense_row_blocks(int n) { return (n + DM - 1) / DM; }

void launch_conv_adj(const ConvArgs& a, int sm_count, cudaStream_t st) {
    static PerDeviceOnce once;
    const size_t smem_small = (size_t)WARPS * WN_SMALL * XS * sizeof(float), smem_big = (size_t)12 * WN_BIG * XS * sizeof(float);
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_adj<WN_SMALL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small));
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_adj<WN_BIG, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
    });
    const ConvGeom g = conv_geom(a.n_tiles, a.wn, sm_count, a.n_chunks);
    if (a.wn == WN_BIG) k_conv_adj<WN_BIG, 12><<<g.blocks, 12 * 32, smem_big, st>>>(a);
    else k_conv_adj<WN_SMALL, WARPS><<<g.blocks, TPB, smem_small, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

void launch_gin(const GinArgs& a, int sm_count, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_gin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gin_smem(false)));
        TGNN_CUDA(cudaFuncSetAttribute(k_gin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gin_smem(true)));
    });
    if (gin_small(a.n_own)) {                    // small graphs: one CTA of four warps per 16-node chunk (latency layout)
        const int chunks = (a.n_own + CH - 1) / CH;
        const size_t sm_s = (size_t)((a.hmlp ? GIN_WFLOATS_H : GIN_WFLOATS) + CH * XS) * sizeof(float);
        if (a.hmlp) k_gin_s<true><<<chunks, GIN_S_THREADS, sm_s, st>>>(a);
        else k_gin_s<false><<<chunks, GIN_S_THREADS, sm_s, st>>>(a);
        TGNN_CUDA(cudaGetLastError());
        return;
    }
    if (a.hmlp) k_gin<true><<<gin_blocks(a.n_own, sm_count), TPB, gin_smem(true), st>>>(a);
    else k_gin<false><<<gin_blocks(a.n_own, sm_count), TPB, gin_smem(false), st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

void launch_node_mask(const Graph& g, const uint8_t* keep, float* inv_deg_masked, int* counters3, double* count, cudaStream_t st) {
    TGNN_CUDA(cudaMemsetAsync(counters3, 0, 3 * sizeof(int), st));
    if (g.n_tiles > 0) k_mask_adj<<<g.n_tiles, 128, 0, st>>>(g.cptr.as<int>(), g.csrc.as<int>(), g.cdst.as<uint8_t>(), g.wn, (int)g.n_own, keep,
                                                            inv_deg_masked, counters3);
    k_mask_count<<<(int)((g.n_own + 255) / 256), 256, 0, st>>>(g.col_ptr.as<int>(), g.col_src.as<int>(), (int)g.n_own, keep, counters3);
    k_mask_finish<<<1, 1, 0, st>>>(counters3, count);
    TGNN_CUDA(cudaGetLastError());
}

void launch_frag_pack_h16(const float* w_kn, int K, int N, int nmap, float* out, int* flag, cudaStream_t st) {
    k_frag_pack_h16<<<(K * N + 255) / 256, 256, 0, st>>>(w_kn, K, N, nmap, reinterpret_cast<__half*>(out), flag);
    TGNN_CUDA(cudaGetLastError());
}

void launch_combine(const float* pre1, const float* coef1, const float* pre2, const float* coef2,
                    const float* residual, float* out, uint4* xh, int* flag, float* g2out, int64_t n_own, cudaStream_t st,
                    const CombineFin* fin, const uint8_t* mask) {
    int64_t n4 = n_own * (F / 4);
    const int tpb = fin ? 1024 : 256;       // the BatchNorm-finishing prologue wants many loads in flight (see the kernel)
    int blocks = (int)std::min<int64_t>((n4 + tpb - 1) / tpb, 148 * 16);
    if (blocks < 1) blocks = 1;
    auto kern = fin ? k_combine<true> : k_combine<false>;
    kern<<<blocks, tpb, 0, st>>>(reinterpret_cast<const float4*>(pre1), coef1,
                                      reinterpret_cast<const float4*>(pre2), coef2,
                                      reinterpret_cast<const float4*>(residual), reinterpret_cast<float4*>(out), xh, flag,
                                      reinterpret_cast<float4*>(g2out), n4, fin ? *fin : CombineFin{}, mask);
    TGNN_CUDA(cudaGetLastError());
}

void launch_init(const InitArgs& a, const InitW1& w1, int mode, int sm_count, cudaStream_t st) {
    int blocks = init_blocks(a.n_own, sm_count);
    if (a.d_x <= INIT_MAX_DX) {
        if (mode == 0) k_init<0><<<blocks, TPB, 0, st>>>(a, w1);
        else if (mode == 1) k_init<1><<<blocks, TPB, 0, st>>>(a, w1);
        else k_init<2><<<blocks, TPB, 0, st>>>(a, w1);
    } else {
        if (mode == 0) k_init_wide<0><<<blocks, TPB, 0, st>>>(a);
        else if (mode == 1) k_init_wide<1><<<blocks, TPB, 0, st>>>(a);
        else k_init_wide<2><<<blocks, TPB, 0, st>>>(a);
    }
    TGNN_CUDA(cudaGetLastError());
}

void launch_dense(const DenseArgs& a, cudaStream_t st) {
    TGNN_CHECK(a.K % DK == 0, "dense stage: K must be a multiple of 32");
    dim3 grid(dense_row_blocks(a.n), 1);
    if (a.n_out % 64 == 0) { grid.y = a.n_out / 64; k_dense<64><<<grid, 256, 0, st>>>(a); }
    else if (a.n_out % 32 == 0) { grid.y = a.n_out / 32; k_dense<32><<<grid, 256, 0, st>>>(a); }
    else TGNN_CHECK(false, "dense stage: n_out must be a multiple of 32");
    TGNN_CUDA(cudaGetLastError());
}

void launch_score(const float* a3, const float* coef, const float* w, float b, float* out, int64_t n, cudaStream_t st, const uint8_t* mask,
                  const BnFin* fin) {
    int blocks = (int)std::min<int64_t>((n + 63) / 64, fin ? 148 : 148 * 8);      // 8 warps x 8 rows per step (every CTA repeats the BatchNorm-finishing prologue)
    if (blocks < 1) blocks = 1;
    k_score<<<blocks, 256, 0, st>>>(a3, coef, w, b, out, n, mask, fin ? *fin : BnFin{});
    TGNN_CUDA(cudaGetLastError());
}

void launch_bn_reduce(const double* part, int n_part, int C, double* sums_out, cudaStream_t st) {
    k_bn_reduce<<<2 * C, 256, 0, st>>>(part, n_part, 2 * C, sums_out);
    TGNN_CUDA(cudaGetLastError());
}
void launch_bn_coef(const double* sums, double count, const float* gamma, const float* beta, float* coef_out, int C,
                    cudaStream_t st) {
    k_bn_coef<<<(C + 63) / 64, 64, 0, st>>>(sums, count, gamma, beta, coef_out, C);
    TGNN_CUDA(cudaGetLastError());
}
void launch_bn_finish(const BnFinishArgs& a, int n_bn, cudaStream_t st) {
    k_bn_finish<<<n_bn * 2 * a.C, 256, 0, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}
void launch_bn_finish_x(const BnFinishArgs& a, int n_bn, const PeerPtrs& p, unsigned epoch, cudaStream_t st) {
    TGNN_CHECK((size_t)n_bn * 2 * a.C <= PX_BN_SLOT, "internal: BatchNorm exchange slot too small");
    k_bn_finish_x<<<n_bn * 2 * a.C, 256, 0, st>>>(a, p, epoch);
    TGNN_CUDA(cudaGetLastError());
}
void launch_halo_push(const float* a, const float* b, const int* rows, int n_send, int64_t halo_slot, const PeerPtrs& p,
                      unsigned epoch, unsigned* ticket, const uint8_t* send_mask, cudaStream_t st) {
    const int64_t n = (int64_t)n_send * 16;
    const int blocks = (int)std::max<int64_t>(1, (n + 255) / 256);           // at least one block: the flags must go up
    k_halo_push<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), rows, n_send,
                                        halo_slot, p, epoch, ticket, send_mask);
    TGNN_CUDA(cudaGetLastError());
}
void launch_halo_unpack_x(const PeerPtrs& p, unsigned epoch, int64_t halo_slot, int64_t n_own, float* a, float* b,
                          uint4* xh, int* flag, const uint8_t* used, unsigned need_from, cudaStream_t st) {
    const int64_t n = (int64_t)p.world * halo_slot * 16;
    const int blocks = (int)std::max<int64_t>(1, (n + 255) / 256);
    k_halo_unpack_x<<<blocks, 256, 0, st>>>(p, epoch, halo_slot, n_own, reinterpret_cast<float4*>(a), reinterpret_cast<float4*>(b), xh, flag, used, need_from);
    TGNN_CUDA(cudaGetLastError());
}
// used[src - n_own] = 1 for every edge source in the mirrored range
__global__ void k_mark_halo(const int64_t* __restrict__ src, int64_t e, int64_t n_own, int64_t n_rows, uint8_t* __restrict__ used) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    const long long s = src[i];
    if (s >= n_own && s < n_rows) used[s - n_own] = 1;
}
void launch_mark_halo(const int64_t* src, int64_t e, int64_t n_own, int64_t n_rows, uint8_t* used, cudaStream_t st) {
    if (e <= 0) return;
    k_mark_halo<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(src, e, n_own, n_rows, used);
    TGNN_CUDA(cudaGetLastError());
}
void launch_bn_coef_eval(const float* rmean, const float* rvar, const float* gamma, const float* beta, float* coef_out,
                         int C, cudaStream_t st) {
    k_bn_coef_eval<<<(C + 63) / 64, 64, 0, st>>>(rmean, rvar, gamma, beta, coef_out, C);
    TGNN_CUDA(cudaGetLastError());
}

void launch_frag_pack(const float* w_kn, int K, int N, int kmap, int nmap, float* out, cudaStream_t st) {
    k_frag_pack<<<(K * N + 255) / 256, 256, 0, st>>>(w_kn, K, N, kmap, nmap, out);
    TGNN_CUDA(cudaGetLastError());
}

void launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st) {
    k_transpose<<<(rows * cols + 255) / 256, 256, 0, st>>>(in, out, rows, cols);
    TGNN_CUDA(cudaGetLastError());
}

void launch_halo_pack(const float* a, const float* b, const int* rows, int n_send, float* sendbuf, cudaStream_t st) {
    if (n_send <= 0) return;
    int64_t n = (int64_t)n_send * 16;
    k_halo_pack<<<(int)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a),
                                                        reinterpret_cast<const float4*>(b), rows, n_send,
                                                        reinterpret_cast<float4*>(sendbuf));
    TGNN_CUDA(cudaGetLastError());
}
void launch_halo_unpack(const float* recv, int world, int rank, int64_t halo_slot, int64_t n_own, float* a, float* b,
                        uint4* xh, int* flag, cudaStream_t st) {
    int64_t n = (int64_t)world * halo_slot * 16;
    if (n <= 0) return;
    k_halo_unpack<<<(int)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(recv), world, rank, halo_slot,
                                                          n_own, reinterpret_cast<float4*>(a), reinterpret_cast<float4*>(b), xh, flag);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
