// GINConv node MLP (32 -> 32 -> 64 -> 32, sigmoid after every layer; graph_networks/layers/coll_conv.py:17-18,25 of the
// reference; PyG GINConv) on a 16-node chunk whose neighbour sums sit in shared memory, followed by LeakyReLU, the store
// of pre2 and the BatchNorm partial sums.  Shared by k_gin (kernels.cu: per-lane global gathers) and k_gin_w (gin_w.cu:
// neighbour rows staged in shared-memory windows by TMA bulk copies).
#pragma once

#include "conv_adj_body.cuh"
#include "hsplit.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace ginx {

using namespace tfx;

__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.0f / (1.0f + expf(-v)); }
// same to within 2 ulp, in 7 instructions instead of 13: exp(-v) = 2^(-v log2 e) on the SFU (ex2.approx: relative error 2^-22;
// the rounding of the product adds |v| 4e-8), reciprocal estimate + one Newton step instead of the IEEE division.  The MLP's
// 128 sigmoids per node are a third of its instructions.  (the clamp keeps 1 + e finite: sigmoid(-80) = 1.8e-35 is below
// every tolerance here)
__device__ __forceinline__ float sigmoidf_nr(float v) {
#ifdef TGNN_SIGMOID_ACC
    { const float x = 1.0f + expf(-fmaxf(v, -80.f)); float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return fmaf(r, fmaf(-x, r, 1.0f), r); }
#endif
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaxf(v, -80.f) * -1.4426950408889634f));
    const float x = 1.0f + e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}


// The MLP's OUTPUT sigmoid feeds BatchNorm directly (eval-mode BatchNorm of the shipped checkpoints amplifies its rounding by
// up to ~300: collapsed running variances), so it keeps the accurate exponential; the two inner layers'
// sigmoids (96 of 128 per node) go through a 32- / 64-term dot product first and use the SFU form.
__device__ __forceinline__ float sigmoidf_out(float v) {
    const float x = 1.0f + expf(-fmaxf(v, -80.f));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

constexpr int GIN_W1 = 2048, GIN_W2 = 4096, GIN_W3 = 4096;                 // floats (hi + lo)
constexpr int GIN_WFLOATS = GIN_W1 + GIN_W2 + GIN_W3 + 128;                // + b1[32] b2[64] b3[32]
constexpr int GIN_W2H = 2048, GIN_W3H = 2048;                              // fp16 hi|lo tables of layers 2, 3 (float-sized words)
constexpr int GIN_WFLOATS_H = GIN_W1 + GIN_W2H + GIN_W3H + 128;
static_assert(TG_GIN_WFLOATS == GIN_WFLOATS + GIN_W2H + GIN_W3H, "gin weight buffer layout");

// weight views of the shared-memory copy made by gin_load_weights
template <bool HMLP>
struct GinW {
    const float4 *W1, *W2, *W3;
    const float *b1, *b2, *b3;
    __device__ __forceinline__ explicit GinW(const float* smem) {
        constexpr int WF = HMLP ? GIN_WFLOATS_H : GIN_WFLOATS;
        W1 = reinterpret_cast<const float4*>(smem);
        W2 = reinterpret_cast<const float4*>(smem + GIN_W1);
        W3 = reinterpret_cast<const float4*>(smem + GIN_W1 + (HMLP ? GIN_W2H : GIN_W2));
        b1 = smem + WF - 128; b2 = b1 + 32; b3 = b2 + 64;
    }
};

// weight buffer (global): [W1 | W2 | W3 | biases] (3xTF32 tables) then [W2h | W3h] (fp16 hi|lo tables) -> shared memory
template <bool HMLP>
__device__ __forceinline__ void gin_load_weights(float* smem, const float* __restrict__ wfrag, int tid, int nthreads) {
    if (HMLP) {
        const float4* src = reinterpret_cast<const float4*>(wfrag);
        float4* dst = reinterpret_cast<float4*>(smem);
        for (int i = tid; i < GIN_W1 / 4; i += nthreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < (GIN_W2H + GIN_W3H) / 4; i += nthreads) dst[GIN_W1 / 4 + i] = __ldg(src + GIN_WFLOATS / 4 + i);
        for (int i = tid; i < 128 / 4; i += nthreads) dst[(GIN_W1 + GIN_W2H + GIN_W3H) / 4 + i] = __ldg(src + (GIN_W1 + GIN_W2 + GIN_W3) / 4 + i);
    } else {
        for (int i = tid; i < GIN_WFLOATS / 4; i += nthreads)
            reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(wfrag) + i);
    }
}

// A fragments of layer 1 for a 16-node chunk whose neighbour sums (+ self term) sit in shared memory as xs[16][XS]
// (natural K order): a1[ks] = {row g, row g+8} x {col 8ks+t, 8ks+t+4}.  Once they are in registers the chunk's
// shared-memory rows may be overwritten.
__device__ __forceinline__ void gin_load_a1(const float* xs, int lane, float (&a1)[4][4]) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a1[ks][0] = xs[g * XS + 8 * ks + t]; a1[ks][1] = xs[(g + 8) * XS + 8 * ks + t];
        a1[ks][2] = xs[g * XS + 8 * ks + t + 4]; a1[ks][3] = xs[(g + 8) * XS + 8 * ks + t + 4];
    }
}

// HMLP: layers 2 and 3 on fp16-split operands (their inputs are sigmoid outputs in (0, 1)); layer 1 sees unbounded sums
// and stays on 3xTF32.
template <bool HMLP>
__device__ __forceinline__ void gin_mlp_chunk(const float (&a1)[4][4], const GinW<HMLP>& Wt, int node0, int n_own, float* __restrict__ out,
                                              double (&s1)[8], double (&s2)[8], int lane, const uint8_t* __restrict__ mask = nullptr) {
    const float4 *W1 = Wt.W1, *W2 = Wt.W2, *W3 = Wt.W3;
    const float *b1 = Wt.b1, *b2 = Wt.b2, *b3 = Wt.b3;
    const int g = lane >> 2, t = lane & 3;
    // ---- layer 1: 32 -> 32 (natural K order) -------------------------------------------------------
    float c1[4][4] = {};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const float (&av)[4] = a1[ks];
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            mma3(c1[2 * j], c1[2 * j + 1], ah, al, W1[((ks * 2 + 0) * 2 + j) * 32 + lane], W1[((ks * 2 + 1) * 2 + j) * 32 + lane]);
    }
    __syncwarp();
    float c3[4][4];
    if (HMLP) {
        const uint4* W2h = reinterpret_cast<const uint4*>(W2);
        const uint4* W3h = reinterpret_cast<const uint4*>(W3);
        // ---- layer 2: 32 -> 64; A fragments = fp16 split of sigmoid(c1 + b1): n-tiles 2ks, 2ks+1 -> k16 step ks
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int nt = 2 * ks + hh;
                const float bA = b1[8 * nt + 2 * t], bB = b1[8 * nt + 2 * t + 1];
                split_h2(sigmoidf_nr(c1[nt][0] + bA), sigmoidf_nr(c1[nt][1] + bB), ah[ks][2 * hh], al[ks][2 * hh]);
                split_h2(sigmoidf_nr(c1[nt][2] + bA), sigmoidf_nr(c1[nt][3] + bB), ah[ks][2 * hh + 1], al[ks][2 * hh + 1]);
            }
        float c2[8][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float sm[2][4] = {}, mn[2][2][4] = {};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint4 bh = W2h[((ks * 2 + 0) * 4 + j) * 32 + lane], bl = W2h[((ks * 2 + 1) * 4 + j) * 32 + lane];
                mma_f16(sm[0], al[ks][0], al[ks][1], al[ks][2], al[ks][3], bh.x, bh.y);
                mma_f16(sm[1], al[ks][0], al[ks][1], al[ks][2], al[ks][3], bh.z, bh.w);
                mma_f16(sm[0], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bl.x, bl.y);
                mma_f16(sm[1], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bl.z, bl.w);
                mma_f16(mn[ks][0], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bh.x, bh.y);
                mma_f16(mn[ks][1], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bh.z, bh.w);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) f4_fma_add(c2[2 * j + u], sm[u], 1.0f / 2048.f, mn[0][u], mn[1][u]);
        }
        // ---- layer 3: 64 -> 32 (output channels 8t..8t+7 per lane); A = fp16 split of sigmoid(c2 + b2) ----------
        uint32_t a3h[4][4], a3l[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int nt = 2 * ks + hh;
                const float bA = b2[8 * nt + 2 * t], bB = b2[8 * nt + 2 * t + 1];
                split_h2(sigmoidf_nr(c2[nt][0] + bA), sigmoidf_nr(c2[nt][1] + bB), a3h[ks][2 * hh], a3l[ks][2 * hh]);
                split_h2(sigmoidf_nr(c2[nt][2] + bA), sigmoidf_nr(c2[nt][3] + bB), a3h[ks][2 * hh + 1], a3l[ks][2 * hh + 1]);
            }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float sm[2][4] = {}, acc[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint4 bh = W3h[((ks * 2 + 0) * 2 + j) * 32 + lane], bl = W3h[((ks * 2 + 1) * 2 + j) * 32 + lane];
                float mn[2][4] = {};
                mma_f16(sm[0], a3l[ks][0], a3l[ks][1], a3l[ks][2], a3l[ks][3], bh.x, bh.y);
                mma_f16(sm[1], a3l[ks][0], a3l[ks][1], a3l[ks][2], a3l[ks][3], bh.z, bh.w);
                mma_f16(sm[0], a3h[ks][0], a3h[ks][1], a3h[ks][2], a3h[ks][3], bl.x, bl.y);
                mma_f16(sm[1], a3h[ks][0], a3h[ks][1], a3h[ks][2], a3h[ks][3], bl.z, bl.w);
                mma_f16(mn[0], a3h[ks][0], a3h[ks][1], a3h[ks][2], a3h[ks][3], bh.x, bh.y);
                mma_f16(mn[1], a3h[ks][0], a3h[ks][1], a3h[ks][2], a3h[ks][3], bh.z, bh.w);
#pragma unroll
                for (int u = 0; u < 2; ++u) {                                     // IEEE adds between the k16 steps (FADD2)
                    const float2 a0 = f2add(make_float2(acc[u][0], acc[u][1]), make_float2(mn[u][0], mn[u][1]));
                    const float2 a1 = f2add(make_float2(acc[u][2], acc[u][3]), make_float2(mn[u][2], mn[u][3]));
                    acc[u][0] = a0.x; acc[u][1] = a0.y; acc[u][2] = a1.x; acc[u][3] = a1.y;
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float2 kk = make_float2(1.0f / 2048.f, 1.0f / 2048.f);
                const float2 r0 = f2fma(make_float2(sm[u][0], sm[u][1]), kk, make_float2(acc[u][0], acc[u][1]));
                const float2 r1 = f2fma(make_float2(sm[u][2], sm[u][3]), kk, make_float2(acc[u][2], acc[u][3]));
                c3[2 * j + u][0] = r0.x; c3[2 * j + u][1] = r0.y; c3[2 * j + u][2] = r1.x; c3[2 * j + u][3] = r1.y;
            }
        }
    } else {
    // ---- layer 2: 32 -> 64, A = sigmoid(c1 + b1) straight from the C fragments ------------------
    float c2[8][4] = {};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const float bA = b1[8 * ks + 2 * t], bB = b1[8 * ks + 2 * t + 1];
        float av[4] = {sigmoidf_nr(c1[ks][0] + bA), sigmoidf_nr(c1[ks][2] + bA),
                       sigmoidf_nr(c1[ks][1] + bB), sigmoidf_nr(c1[ks][3] + bB)};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            mma3(c2[2 * j], c2[2 * j + 1], ah, al, W2[((ks * 2 + 0) * 4 + j) * 32 + lane], W2[((ks * 2 + 1) * 4 + j) * 32 + lane]);
    }
    // ---- layer 3: 64 -> 32 (output channels 8t..8t+7 per lane) ----------------------------------
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) c3[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const float bA = b2[8 * ks + 2 * t], bB = b2[8 * ks + 2 * t + 1];
        float av[4] = {sigmoidf_nr(c2[ks][0] + bA), sigmoidf_nr(c2[ks][2] + bA),
                       sigmoidf_nr(c2[ks][1] + bB), sigmoidf_nr(c2[ks][3] + bB)};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            mma3(c3[2 * j], c3[2 * j + 1], ah, al, W3[((ks * 2 + 0) * 2 + j) * 32 + lane], W3[((ks * 2 + 1) * 2 + j) * 32 + lane]);
    }
    }
    // ---- sigmoid, LeakyReLU, store, statistics ------------------------------------------------
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int node = node0 + g + 8 * half;
        float o[8];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                o[2 * nt + e] = leaky(sigmoidf_out(c3[nt][2 * half + e] + b3[8 * t + 2 * nt + e]));
        if (node < n_own) {
            if (!row_kept(mask, node)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = 0.f;
            }
            float4* dst = reinterpret_cast<float4*>(out + (size_t)node * F + 8 * t);
            dst[0] = make_float4(o[0], o[1], o[2], o[3]);
            dst[1] = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] += (double)o[j]; s2[j] += (double)o[j] * (double)o[j]; }
        }
    }
}

}  // namespace ginx
}  // namespace tgnn
