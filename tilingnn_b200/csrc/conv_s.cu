// Adjacency branch on the 5th-generation tensor cores ("S" formulation).
//
// NNConv(mean) is  out_i = 1/deg_i * sum_e x[src_e] W_{type(e)} + x_i root + bias.  Grouping a destination's
// in-edges by type gives  sum_t ( sum_{e of type t} x[src_e] ) W_t : for a tile of 128 destinations and one
// type t, the rows  S_t[r] = sum of the source rows of dst r's type-t in-edges  (zero if none) form a dense
// [128 x 32] operand, and  D[128 x 32] += S_t W_t  is exactly one tcgen05 tile product.  The accumulator stays
// in TMEM over all types present in the tile (one PASS per type) -- no shared-memory scatter of messages, no
// atomics, cost per (destination, type) instead of per edge, and M = 128 is what tcgen05 wants.  (The edge-chunk
// mma.sync kernel in kernels.cu remains the path for graphs with many edge types, where K passes do not pay.)
//
// CTA = one tile of 128 destinations, 2 CTAs per SM.
//   producer group g (4 warps, passes p = g, g+2, ...): gather + pre-sum the source rows into registers (the
//     pass's row offsets, source indices and weight tiles were prefetched one pass ahead with cp.async), then
//     wait for its smem stage, hi/lo TF32 split, swizzled store, fence, mbarrier arrive
//   MMA warp: per pass 12 x tcgen05.mma.kind::tf32 (M=128, N=32; 3xTF32), commit -> stage empty
//   the root term x_i root is one more pass into a second accumulator (columns 32..63)
//   epilogue (warps 0-3): tcgen05.ld, * 1/deg + root + bias, LeakyReLU, store, BatchNorm partial sums (fp64)
#include <algorithm>

#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {
using namespace tc;

constexpr int SA_TILE = 16384;                 // bytes of one A tile (hi or lo): 128 rows x 128 B
constexpr int SB_TILE = 4096;                  // bytes of one B tile (hi or lo): 32 rows x 128 B
constexpr int IDX_CAP = 512;                   // staged source indices per pass
constexpr int CS_THREADS = 9 * 32;             // 2 producer groups x 4 warps + MMA warp
constexpr int OFF_A = 0;                                   // [2 stages][hi|lo]
constexpr int OFF_B = OFF_A + 4 * SA_TILE;                 // [2 groups][2 slots][hi|lo]
constexpr int OFF_IDX = OFF_B + 8 * SB_TILE;               // [2 groups][2 slots][IDX_CAP] int
constexpr int OFF_OFFB = OFF_IDX + 4 * IDX_CAP * 4;        // [2 groups][2 slots][S_OFF_STRIDE] uint16
constexpr int OFF_SCAL = OFF_OFFB + 4 * S_OFF_STRIDE * 2;  // ptype[128], pbase[129]
constexpr int CS_SMEM = OFF_SCAL + (128 + 132) * 4 + 1024;

__device__ __forceinline__ float4 ld_row4(const float* base, int row, int q) {
    return __ldg(reinterpret_cast<const float4*>(base + (size_t)row * F) + q);
}

struct ConvSArgs {
    const float* xin;            // [n_rows][32]
    const float* tabS;           // [K+1][hi|lo][32 n][32 k]  W_t transposed (entry K = root transposed)
    int n_types;
    const int* pptr; const int* ptype; const int* pbase; const unsigned short* off; const int* ssrc;
    const float* inv_deg; const float* bias;
    float* out; double* part; int* error_flag;
    int n_own, n_tiles;
};

__global__ void __launch_bounds__(CS_THREADS, 2)
k_conv_s(ConvSArgs A) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[5];          // full[2], empty[2], acc
    __shared__ uint32_t tmem_base_smem;
    __shared__ int timeout_flag;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    int* s_ptype = reinterpret_cast<int*>(smem + OFF_SCAL);
    int* s_pbase = s_ptype + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[2]), bar_acc = smem_u32(&bars[4]);
    const int tile = blockIdx.x;
    const int p0 = __ldg(A.pptr + tile), np = __ldg(A.pptr + tile + 1) - p0;   // typed passes of this tile

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(bar_full + 8 * s, 4); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_acc, 1);
        timeout_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i <= np && i < 129; i += CS_THREADS) {
        if (i < np) s_ptype[i] = __ldg(A.ptype + p0 + i);
        s_pbase[i] = __ldg(A.pbase + p0 + i);
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < 8) {
        // ===================== producers =====================
        const int g = warp >> 2, gt = tid & 127, c = gt & 7, rbase = gt >> 3;
        const uint32_t bar_id = 2 + g;
        auto prefetch = [&](int q, int slot) {
            if (q > np) return;
            const int type = q < np ? s_ptype[q] : A.n_types;
            const uint32_t bdst = smem_base + OFF_B + (uint32_t)((g * 2 + slot) * 2) * SB_TILE;
            for (int i = gt; i < 512; i += 128) {
                const int hl = i >> 8, n = (i >> 3) & 31, cc = i & 7;
                const float* src = A.tabS + ((size_t)(type * 2 + hl) * 32 + n) * 32 + 4 * cc;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(bdst + hl * SB_TILE + sw128_off(n, cc)), "l"(src) : "memory");
            }
            if (q < np) {
                const uint32_t odst = smem_base + OFF_OFFB + (uint32_t)(g * 2 + slot) * S_OFF_STRIDE * 2;
                if (gt < 17) {
                    const unsigned short* src = A.off + (size_t)(p0 + q) * S_OFF_STRIDE + gt * 8;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(odst + gt * 16), "l"(src) : "memory");
                }
                const int base = s_pbase[q], len = s_pbase[q + 1] - base;
                if (len <= IDX_CAP) {
                    const uint32_t idst = smem_base + OFF_IDX + (uint32_t)(g * 2 + slot) * IDX_CAP * 4;
                    for (int i = gt; i < len; i += 128)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(idst + i * 4), "l"(A.ssrc + base + i) : "memory");
                }
            }
        };
        prefetch(g, 0);
        bool ok = true;
        for (int q = g, k = 0; q <= np && ok; q += 2, ++k) {
            const int slot = k & 1;
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            // ---- gather + pre-sum into registers (overlaps the MMAs of earlier passes) ----
            float4 v[8];
            if (q < np) {
                const unsigned short* ob = reinterpret_cast<const unsigned short*>(smem + OFF_OFFB) + (g * 2 + slot) * S_OFF_STRIDE;
                const int* ib = reinterpret_cast<const int*>(smem + OFF_IDX) + (g * 2 + slot) * IDX_CAP;
                const int base = s_pbase[q];
                const bool staged = (s_pbase[q + 1] - base) <= IDX_CAP;
                int es[8], ee[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { es[j] = ob[rbase + 16 * j]; ee[j] = ob[rbase + 16 * j + 1]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (es[j] < ee[j]) {
                        const int idx = staged ? ib[es[j]] : __ldg(A.ssrc + base + es[j]);
                        v[j] = ld_row4(A.xin, idx, c);
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    for (int t = es[j] + 1; t < ee[j]; ++t) {
                        const int idx = staged ? ib[t] : __ldg(A.ssrc + base + t);
                        const float4 w = ld_row4(A.xin, idx, c);
                        v[j].x += w.x; v[j].y += w.y; v[j].z += w.z; v[j].w += w.w;
                    }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int node = tile * S_BM + rbase + 16 * j;
                    v[j] = node < A.n_own ? ld_row4(A.xin, node, c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // ---- stage g is free once the MMAs of pass q-2 have read it ----
            if (!mbar_wait(bar_empty + 8 * g, (k & 1) ^ 1)) { timeout_flag = 1; ok = false; break; }
            prefetch(q + 2, slot ^ 1);
            uint8_t* sa_hi = smem + OFF_A + (g * 2) * SA_TILE;
            uint8_t* sa_lo = sa_hi + SA_TILE;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 hi, lo;
                hi.x = tf32_rna(v[j].x); lo.x = tf32_rna(v[j].x - __uint_as_float(hi.x));
                hi.y = tf32_rna(v[j].y); lo.y = tf32_rna(v[j].y - __uint_as_float(hi.y));
                hi.z = tf32_rna(v[j].z); lo.z = tf32_rna(v[j].z - __uint_as_float(hi.z));
                hi.w = tf32_rna(v[j].w); lo.w = tf32_rna(v[j].w - __uint_as_float(hi.w));
                const uint32_t o = sw128_off(rbase + 16 * j, c);
                *reinterpret_cast<uint4*>(sa_hi + o) = hi;
                *reinterpret_cast<uint4*>(sa_lo + o) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * g);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (lane == 0) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC = umma_idesc_tf32(32);
        for (int q = 0; q <= np; ++q) {
            const int g = q & 1, k = q >> 1, slot = k & 1;
            if (!mbar_wait(bar_full + 8 * g, k & 1)) { timeout_flag = 1; break; }
            tc_fence_after();
            const uint32_t a_hi = smem_base + OFF_A + (g * 2) * SA_TILE, a_lo = a_hi + SA_TILE;
            const uint32_t b_hi = smem_base + OFF_B + (uint32_t)((g * 2 + slot) * 2) * SB_TILE, b_lo = b_hi + SB_TILE;
            const bool root = q == np;
            const uint32_t tmem_d = tmem_base + (root ? 32u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t kb = ks * 32;
                const uint64_t dah = umma_desc_sw128(a_hi + kb), dal = umma_desc_sw128(a_lo + kb);
                const uint64_t dbh = umma_desc_sw128(b_hi + kb), dbl = umma_desc_sw128(b_lo + kb);
                umma_tf32(tmem_d, dal, dbh, IDESC, (ks > 0 || (!root && q > 0)) ? 1u : 0u);
                umma_tf32(tmem_d, dah, dbl, IDESC, 1u);
                umma_tf32(tmem_d, dah, dbh, IDESC, 1u);
            }
            umma_commit(bar_empty + 8 * g);
        }
        umma_commit(bar_acc);
    }

    // ===================== epilogue: warps 0-3, tile row = 32 * warp + lane =====================
    float* scratch = reinterpret_cast<float*>(smem + OFF_A);                    // [4][32*33] floats
    double* red = reinterpret_cast<double*>(smem + OFF_A + 4 * 32 * 33 * 4);    // [4][2][32]
    bool acc_ok = true;
    if (warp < 4) {
        acc_ok = mbar_wait(bar_acc, 0);
        if (!acc_ok) timeout_flag = 1;
        tc_fence_after();
        const int row = tile * S_BM + 32 * warp + lane;
        const bool live = row < A.n_own && acc_ok;
        uint32_t vt[32], vr[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16) + 32u, vr);
        if (np > 0) tmem_ld32(tmem_base + ((uint32_t)(32 * warp) << 16), vt);
        const float idg = live ? __ldg(A.inv_deg + row) : 0.f;
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float t = np > 0 ? __uint_as_float(vt[j]) : 0.f;
            o[j] = leaky(fmaf(t, idg, __uint_as_float(vr[j])) + __ldg(A.bias + j));
        }
        if (live) {
            float4* dst = reinterpret_cast<float4*>(A.out + (size_t)row * F);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
        if (A.part) {
            float* sc = scratch + warp * (32 * 33);
#pragma unroll
            for (int j = 0; j < 32; ++j) sc[lane * 33 + j] = o[j];
            __syncwarp();
            int nv = A.n_own - (tile * S_BM + 32 * warp);
            nv = (!acc_ok || nv < 0) ? 0 : (nv > 32 ? 32 : nv);
            double s1 = 0.0, s2 = 0.0;
            for (int r = 0; r < nv; ++r) { const double x = (double)sc[r * 33 + lane]; s1 += x; s2 += x * x; }
            red[(warp * 2 + 0) * 32 + lane] = s1;
            red[(warp * 2 + 1) * 32 + lane] = s2;
        }
        tc_fence_before();
    }
    __syncthreads();
    if (A.part && tid < 64) {
        const int qq = tid >> 5, cc = tid & 31;
        A.part[(size_t)tile * 64 + tid] = ((red[(0 * 2 + qq) * 32 + cc] + red[(1 * 2 + qq) * 32 + cc]) + red[(2 * 2 + qq) * 32 + cc]) +
                                          red[(3 * 2 + qq) * 32 + cc];
    }
    if (timeout_flag && tid == 0) atomicExch(A.error_flag, 1);
    if (warp == 8) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(64));
    }
}

// per-type edge weights for the S kernel: tabS[t][hi|lo][n][k] = TF32 split of W_t[k][n], W_t evaluated in fp64
__global__ void k_edge_table_s(const float* __restrict__ rows, int d_e, const float* __restrict__ a1, const float* __restrict__ c1,
                               const float* __restrict__ a2, const float* __restrict__ c2, const float* __restrict__ a3,
                               const float* __restrict__ c3, float* __restrict__ tab) {
    __shared__ double h1[32], h2[64];
    const int t = blockIdx.x, tid = threadIdx.x;
    const float* e = rows + (size_t)t * d_e;
    if (tid < 32) {
        double s = (double)c1[tid];
        for (int k = 0; k < d_e; ++k) s += (double)a1[tid * d_e + k] * (double)e[k];
        h1[tid] = 1.0 / (1.0 + exp(-s));
    }
    __syncthreads();
    if (tid < 64) {
        double s = (double)c2[tid];
        for (int k = 0; k < 32; ++k) s += (double)a2[tid * 32 + k] * h1[k];
        h2[tid] = 1.0 / (1.0 + exp(-s));
    }
    __syncthreads();
    float* out = tab + (size_t)t * 2048;
    for (int o = tid; o < F * F; o += 256) {
        double s = (double)c3[o];
        for (int k = 0; k < 64; ++k) s += (double)a3[(size_t)o * 64 + k] * h2[k];
        const double w = 1.0 / (1.0 + exp(-s));
        const int kin = o >> 5, n = o & 31;                   // NNConv: weight.view(-1, in, out)
        const uint32_t hi = tf32_rna((float)w);
        const uint32_t lo = tf32_rna((float)(w - (double)__uint_as_float(hi)));
        out[n * 32 + kin] = __uint_as_float(hi);
        out[1024 + n * 32 + kin] = __uint_as_float(lo);
    }
}
// root [in][out] -> [hi|lo][n = out][k = in]
__global__ void k_root_table_s(const float* __restrict__ root, float* __restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= F * F) return;
    const int kin = o >> 5, n = o & 31;
    const float w = root[o];
    const uint32_t hi = tf32_rna(w);
    out[n * 32 + kin] = __uint_as_float(hi);
    out[1024 + n * 32 + kin] = __uint_as_float(tf32_rna(w - __uint_as_float(hi)));
}

}  // namespace

void launch_edge_table_s(const float* type_rows, int n_types, int d_e, const float* a1, const float* c1, const float* a2,
                         const float* c2, const float* a3, const float* c3, const float* root, float* tab, cudaStream_t st) {
    if (n_types > 0) k_edge_table_s<<<n_types, 256, 0, st>>>(type_rows, d_e, a1, c1, a2, c2, a3, c3, tab);
    k_root_table_s<<<4, 256, 0, st>>>(root, tab + (size_t)n_types * 2048);
    TGNN_CUDA(cudaGetLastError());
}

void launch_conv_s(const ConvArgs& c, const Graph& g, const float* tabS, int* error_flag, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_s, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
        attr = true;
    }
    ConvSArgs a{};
    a.xin = c.xin; a.tabS = tabS; a.n_types = g.n_types;
    a.pptr = g.s_pptr.as<int>(); a.ptype = g.s_ptype.as<int>(); a.pbase = g.s_pbase.as<int>();
    a.off = g.s_off.as<unsigned short>(); a.ssrc = g.s_src.as<int>();
    a.inv_deg = c.inv_deg; a.bias = c.bias; a.out = c.out; a.part = c.part; a.error_flag = error_flag;
    a.n_own = c.n_own; a.n_tiles = g.s_tiles;
    k_conv_s<<<g.s_tiles, CS_THREADS, CS_SMEM, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
