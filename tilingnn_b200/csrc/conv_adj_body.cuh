// 3xTF32 edge-chunk arithmetic of the adjacency branch, shared by k_conv_adj (kernels.cu) and -- as the wide-range
// stand-in taken when a range flag is raised -- by k_conv_h (conv_h.cu); also the mma.sync 3xTF32 building block of k_gin.
#pragma once

#include "bn_fin.cuh"
#include "hsplit.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace tfx {

constexpr int XS = 36;          // padded shared-memory row stride (floats): 144 B keeps float4 alignment
__device__ __forceinline__ float leaky(float v) { return v >= 0.f ? v : v * LEAKY; }

// ------------------------------------------------------------------------------------------------
// Tensor-core building block: mma.sync m16n8k8 TF32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi),
// which keeps fp32-level accuracy (dropped term ~2^-22) -- needed for the 1e-4 parity bar, single
// TF32 (2^-11) is not enough.  The irregular 16-row chunks of this path (16 gathered edges of one
// type; 16 nodes of the GIN MLP) are below tcgen05's minimum M of 64, so they use the warp-level
// mma.sync path; operands are laid out so that NO shared-memory staging of A is needed:
//   * K is permuted so a lane's two float4 loads of a gathered row ARE its A fragments,
//   * N is permuted so a lane ends up with 8 contiguous output channels (float4 RMW / stores),
//   * chained layers use the previous C fragments directly as the next A fragments.
// B fragments come from tables pre-split into hi/lo and stored in fragment order ("frag tables"):
//   float4 index ((ks*2 + hl) * (N/16) + j) * 32 + lane ; float4 = {b0,b1 of n-tile 2j, b0,b1 of n-tile 2j+1}
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rna(x);
    lo = tf32_rna(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 3xTF32 update of one n-tile pair held in a float4 of hi and a float4 of lo B fragments.
// Tensor cores accumulate with round-toward-zero; chaining every MMA into one accumulator builds a
// one-sided error of ~1 ulp per instruction that train-mode BatchNorm cancels but eval-mode BatchNorm
// amplifies.  So each k-step's three products go into a zeroed temporary (small terms first) and are
// added to the running sum with an IEEE round-to-nearest FADD (Ootomo & Yokota's 3xTF32 recipe).
__device__ __forceinline__ void mma3(float (&c0)[4], float (&c1)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                     const float4& bh, const float4& bl) {
    float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
    mma_tf32(t0, alo, __float_as_uint(bh.x), __float_as_uint(bh.y));
    mma_tf32(t0, ahi, __float_as_uint(bl.x), __float_as_uint(bl.y));
    mma_tf32(t0, ahi, __float_as_uint(bh.x), __float_as_uint(bh.y));
    mma_tf32(t1, alo, __float_as_uint(bh.z), __float_as_uint(bh.w));
    mma_tf32(t1, ahi, __float_as_uint(bl.z), __float_as_uint(bl.w));
    mma_tf32(t1, ahi, __float_as_uint(bh.z), __float_as_uint(bh.w));
#pragma unroll
    for (int i = 0; i < 4; ++i) { c0[i] += t0[i]; c1[i] += t1[i]; }
}

__device__ __forceinline__ float4 ld_row4(const float* base, int row, int q) {
    return __ldg(reinterpret_cast<const float4*>(base + (size_t)row * F) + q);
}

// ------------------------------------------------------------------------------------------------
// Adjacency branch: typed NNConv(mean) + root + bias + LeakyReLU, BatchNorm partial sums.
// (graph_networks/layers/edge_conv.py:24-27 of the reference; PyG NNConv semantics.)
// One warp owns a tile of WN destination rows and walks its chunks of 16 same-type edges:
//   gather (2 x LDG.128 per row per lane, straight into A fragments) -> 48 mma.sync (3xTF32) against the
//   type's B fragments held in registers -> accumulate the 16 messages into the warp's private
//   shared-memory tile (float4 read-modify-write, no atomics: destinations are distinct per 8-slot group).
// ------------------------------------------------------------------------------------------------
constexpr int FRAG32 = 2048;    // floats of one 32x32 frag table (hi + lo)

struct BFrag32 { float4 h[4][2], l[4][2]; };

__device__ __forceinline__ void load_bfrag32(BFrag32& b, const float* __restrict__ tab, int lane) {
    const float4* p = reinterpret_cast<const float4*>(tab) + lane;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            b.h[ks][j] = __ldg(p + ((ks * 2 + 0) * 2 + j) * 32);
            b.l[ks][j] = __ldg(p + ((ks * 2 + 1) * 2 + j) * 32);
        }
}

// rows[0..1] = the lane's two float4 of row g, rows[2..3] = of row g+8 (KMAP_GATHER)
__device__ __forceinline__ void chunk_mma32(const float4 (&rows)[4], const BFrag32& b, float (&c)[4][4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const float4 lo4 = rows[ks >> 1], hi4 = rows[2 + (ks >> 1)];
        float av[4];
        av[0] = (ks & 1) ? lo4.z : lo4.x;   // (row g,   k = t)
        av[1] = (ks & 1) ? hi4.z : hi4.x;   // (row g+8, k = t)
        av[2] = (ks & 1) ? lo4.w : lo4.y;   // (row g,   k = t+4)
        av[3] = (ks & 1) ? hi4.w : hi4.y;   // (row g+8, k = t+4)
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
        mma3(c[0], c[1], ah, al, b.h[ks][0], b.l[ks][0]);
        mma3(c[2], c[3], ah, al, b.h[ks][1], b.l[ks][1]);
    }
}

__device__ __forceinline__ void acc_add8(float* row, const float (&c)[4][4], int half) {
    float4* p = reinterpret_cast<float4*>(row);
    const float4 v0 = p[0], v1 = p[1];
    const float2 a = f2add(make_float2(v0.x, v0.y), make_float2(c[0][2 * half], c[0][2 * half + 1]));       // FADD2: the C fragment's register
    const float2 b = f2add(make_float2(v0.z, v0.w), make_float2(c[1][2 * half], c[1][2 * half + 1]));       // pairs are adjacent columns
    const float2 d = f2add(make_float2(v1.x, v1.y), make_float2(c[2][2 * half], c[2][2 * half + 1]));
    const float2 e = f2add(make_float2(v1.z, v1.w), make_float2(c[3][2 * half], c[3][2 * half + 1]));
    p[0] = make_float4(a.x, a.y, b.x, b.y); p[1] = make_float4(d.x, d.y, e.x, e.y);
}

template <int WN, int NW>
__device__ __forceinline__ void conv_adj_body(const ConvArgs& A, float* smem) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* acc = smem + warp * (WN * XS);
    const int g = lane >> 2, t = lane & 3;
    const int gwarp = blockIdx.x * NW + warp, nwarp = gridDim.x * NW;
    double s1 = 0.0, s2 = 0.0;
    const float bias_c = __ldg(A.bias + lane);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    BFrag32 bf;
    int cur_type = -1;

    for (int tile = gwarp; tile < A.n_tiles; tile += nwarp) {
        for (int i = lane; i < WN * XS; i += 32) acc[i] = 0.f;
        const int c0 = __ldg(A.cptr + tile), c1 = __ldg(A.cptr + tile + 1);
        float4 pre[4];
        int psrc = -1, pdst = 0, ptype = 0;
        if (c0 < c1) {
            psrc = __ldg(A.csrc + (size_t)c0 * CH + (lane & 15));
            pdst = __ldg(A.cdst + (size_t)c0 * CH + (lane & 15));
            ptype = __ldg(A.ctype + c0);
            const int sa = __shfl_sync(0xffffffffu, psrc, g), sb = __shfl_sync(0xffffffffu, psrc, g + 8);
            pre[0] = sa >= 0 ? ld_row4(A.xin, sa, t) : zero4; pre[1] = sa >= 0 ? ld_row4(A.xin, sa, 4 + t) : zero4;
            pre[2] = sb >= 0 ? ld_row4(A.xin, sb, t) : zero4; pre[3] = sb >= 0 ? ld_row4(A.xin, sb, 4 + t) : zero4;
        }
        __syncwarp();
        for (int c = c0; c < c1; ++c) {
            const float4 cur[4] = {pre[0], pre[1], pre[2], pre[3]};
            const int csrc = psrc, cdst = pdst, type = ptype;
            if (c + 1 < c1) {
                psrc = __ldg(A.csrc + (size_t)(c + 1) * CH + (lane & 15));
                pdst = __ldg(A.cdst + (size_t)(c + 1) * CH + (lane & 15));
                ptype = __ldg(A.ctype + c + 1);
                const int sa = __shfl_sync(0xffffffffu, psrc, g), sb = __shfl_sync(0xffffffffu, psrc, g + 8);
                pre[0] = sa >= 0 ? ld_row4(A.xin, sa, t) : zero4; pre[1] = sa >= 0 ? ld_row4(A.xin, sa, 4 + t) : zero4;
                pre[2] = sb >= 0 ? ld_row4(A.xin, sb, t) : zero4; pre[3] = sb >= 0 ? ld_row4(A.xin, sb, 4 + t) : zero4;
            }
            if (type != cur_type) { load_bfrag32(bf, A.tabF + (size_t)type * FRAG32, lane); cur_type = type; }
            // the next chunk's type is already known (ptype): pull its 8 KB fragment table towards L1 now, so the
            // reload at the type change does not expose an L2 round trip (64 lines of 128 B, two per lane)
            if (ptype != type && c + 1 < c1) {
                const char* nt = reinterpret_cast<const char*>(A.tabF + (size_t)ptype * FRAG32);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nt + lane * 128));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(nt + (lane + 32) * 128));
            }
            float m[4][4] = {};
            chunk_mma32(cur, bf, m);
            // rows 0..7 (group 0), then rows 8..15 (group 1): destinations are distinct inside a group
            {
                const int s = __shfl_sync(0xffffffffu, csrc, g), d = __shfl_sync(0xffffffffu, cdst, g);
                if (s >= 0) acc_add8(acc + d * XS + 8 * t, m, 0);
            }
            __syncwarp();
            {
                const int s = __shfl_sync(0xffffffffu, csrc, g + 8), d = __shfl_sync(0xffffffffu, cdst, g + 8);
                if (s >= 0) acc_add8(acc + d * XS + 8 * t, m, 1);
            }
            __syncwarp();
        }
        // mean over in-edges
        const int node0 = tile * WN;
        for (int r = 0; r < WN; ++r) {
            int node = node0 + r;
            if (node < A.n_own) acc[r * XS + lane] *= __ldg(A.inv_deg + node);
        }
        __syncwarp();
        // root term: x_i @ root as four 16-row chunks of the tile's own rows (frag table entry n_types)
        if (cur_type != A.n_types) { load_bfrag32(bf, A.tabF + (size_t)A.n_types * FRAG32, lane); cur_type = A.n_types; }
        for (int rc = 0; rc < WN / CH; ++rc) {
            const int na = node0 + rc * CH + g, nb = na + 8;
            float4 cur[4];
            cur[0] = na < A.n_own ? ld_row4(A.xin, na, t) : zero4; cur[1] = na < A.n_own ? ld_row4(A.xin, na, 4 + t) : zero4;
            cur[2] = nb < A.n_own ? ld_row4(A.xin, nb, t) : zero4; cur[3] = nb < A.n_own ? ld_row4(A.xin, nb, 4 + t) : zero4;
            float m[4][4] = {};
            chunk_mma32(cur, bf, m);
            acc_add8(acc + (rc * CH + g) * XS + 8 * t, m, 0);
            acc_add8(acc + (rc * CH + g + 8) * XS + 8 * t, m, 1);
        }
        __syncwarp();
        // bias, LeakyReLU, store, statistics (lane = channel)
        for (int r = 0; r < WN; ++r) {
            int node = node0 + r;
            if (node < A.n_own) {
                float v = leaky(acc[r * XS + lane] + bias_c);
                if (!row_kept(A.mask, node)) v = 0.f;
                A.out[(size_t)node * F + lane] = v;
                s1 += (double)v;
                s2 += (double)v * (double)v;
            }
        }
        __syncwarp();
    }
    if (A.part) block_part_store(A.part, s1, s2, reinterpret_cast<double*>(smem), NW);      // one partial row per CTA
}


}  // namespace tfx
}  // namespace tgnn
