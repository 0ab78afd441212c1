// Adjacency branch on PRE-SPLIT fp16 operands ("H" kernel): typed NNConv(mean) + root + bias + LeakyReLU and the
// BatchNorm partial sums (graph_networks/layers/edge_conv.py:24-27 of the reference; PyG NNConv semantics).
//
// Same edge-chunk formulation as k_conv_adj (kernels.cu): a warp owns 64 destination rows and walks chunks of 16
// same-type edges.  What changes is the arithmetic of the 16x32 . 32x32 products:
//   * every producer of b1 (k_init<2>, k_combine, k_halo_unpack) also writes a SPLIT copy of the row,
//       x = hi + lo * 2^-11,   hi = fp16(x),  lo = fp16((x - hi) * 2^11)          (22 significant bits)
//     laid out so that one LDG.128 of a gathered row IS the lane's A fragments of a k16 step (hi and lo);
//   * the per-type weights are split the same way into fp16 fragment tables (4 KB per type);
//   * x.W ~= hi.Whi + (hi.Wlo + lo.Whi) * 2^-11  on  mma.sync.m16n8k16.f16 (fp32 accumulate): 24 MMAs per chunk
//     instead of the 48 of 3xTF32, and no conversion instructions in the consumer at all.  The dropped lo.lo term
//     is 2^-22 relative -- the same as 3xTF32 (measured: 1.0e-7 vs 1.2e-7 of sum|x||w|).
// fp16 has a narrow exponent range, so the producers raise a per-tensor flag when |x| > 60000 (or NaN); this kernel
// then runs the layer with the 3xTF32 arithmetic of k_conv_adj on the fp32 rows instead (conv_adj_body.cuh).  Values below 2^-14 keep an absolute accuracy of 2^-35, far under fp32 rounding of the O(1) sums.
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <string>

#include "conv_adj_body.cuh"
#include "hsplit.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {

namespace cg = cooperative_groups;

using tfx::XS;                  // padded shared-memory row stride (floats)
constexpr float LO_INV = 1.0f / 2048.f;

using tfx::leaky;
using tfx::acc_add8;

struct BFragH { uint4 h[2][2], l[2][2]; };

__device__ __forceinline__ void load_bfrag_h(BFragH& b, const uint32_t* __restrict__ tab, int lane) {
    const uint4* p = reinterpret_cast<const uint4*>(tab) + lane;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            b.h[ks][j] = __ldg(p + ((ks * 2 + 0) * 2 + j) * 32);
            b.l[ks][j] = __ldg(p + ((ks * 2 + 1) * 2 + j) * 32);
        }
}

// rows[ks] = uint4 (4ks' + t) of row g, rows[2 + ks] = of row g+8:  {hi(c0,c1), hi(c2,c3), lo(c0,c1), lo(c2,c3)}
// Tensor cores accumulate with truncation; the main term of each k16 step gets its own zeroed accumulator and the
// steps are combined with IEEE FADDs (same reasoning as mma3 in kernels.cu); the two small terms share one.
// The A operand of mma.sync.m16n8k16 is a register QUAD {row g k-lo, row g+8 k-lo, row g k-hi, row g+8 k-hi}: it interleaves the
// pieces of two gathered rows, i.e. of two different 256-bit loads.  Left to itself the compiler keeps the rows as loaded
// and re-assembles a quad in front of nearly every HMMA pair (the HMMA results overwrite the assembled copy): ~52 moves per
// chunk, plus 16 for carrying the prefetched rows into the next iteration.  Here the four quads of a chunk
// ([0] hi, k16 step 0; [1] lo, step 0; [2] hi, step 1; [3] lo, step 1) are the ONLY live form of its rows: they are assembled
// once, when the prefetched rows are taken over at the end of the previous iteration (16 moves in all).
struct AQuads { uint32_t q[4][4]; };
// rows[ks] = uint4 (4ks' + t) of row g, rows[2 + ks] = of row g+8:  {hi(c0,c1), hi(c2,c3), lo(c0,c1), lo(c2,c3)}
__device__ __forceinline__ void make_quads(const uint4 (&rows)[4], AQuads& a) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const uint4 ra = rows[ks], rb = rows[2 + ks];
        a.q[2 * ks][0] = ra.x; a.q[2 * ks][1] = rb.x; a.q[2 * ks][2] = ra.y; a.q[2 * ks][3] = rb.y;
        a.q[2 * ks + 1][0] = ra.z; a.q[2 * ks + 1][1] = rb.z; a.q[2 * ks + 1][2] = ra.w; a.q[2 * ks + 1][3] = rb.w;
    }
}
// Tensor cores accumulate with truncation; the main term of each k16 step gets its own zeroed accumulator and the
// steps are combined with IEEE FADDs (same reasoning as mma3 in kernels.cu); the two small terms share one.
__device__ __forceinline__ void chunk_mma_q(const AQuads& a, const BFragH& b, float (&m)[4][4]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float sm[2][4] = {}, mn[2][2][4] = {};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t (&hi)[4] = a.q[2 * ks];
            const uint32_t (&lo)[4] = a.q[2 * ks + 1];
            const uint4 bh = b.h[ks][j], bl = b.l[ks][j];
            // issue order: the two updates of sm[u] are four instructions apart (an HMMA's result is ready after ~4
            // issue slots of the tensor pipe), the independent main products sit between them
            mma_f16(sm[0], lo[0], lo[1], lo[2], lo[3], bh.x, bh.y);          // lo . Whi
            mma_f16(sm[1], lo[0], lo[1], lo[2], lo[3], bh.z, bh.w);
            mma_f16(mn[ks][0], hi[0], hi[1], hi[2], hi[3], bh.x, bh.y);      // hi . Whi
            mma_f16(mn[ks][1], hi[0], hi[1], hi[2], hi[3], bh.z, bh.w);
            mma_f16(sm[0], hi[0], hi[1], hi[2], hi[3], bl.x, bl.y);          // hi . Wlo
            mma_f16(sm[1], hi[0], hi[1], hi[2], hi[3], bl.z, bl.w);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) f4_fma_add(m[2 * j + u], sm[u], LO_INV, mn[0][u], mn[1][u]);      // FADD2 + FFMA2
    }
}
__device__ __forceinline__ void chunk_mma_h(const uint4 (&rows)[4], const BFragH& b, float (&m)[4][4]) {
    AQuads a;
    make_quads(rows, a);
    chunk_mma_q(a, b, m);
}

// shared-memory form of acc_add8 (conv_adj_body.cuh) on a 32-bit shared address: the hot loop then forms a scatter address
// with ONE multiply-add (lane base + destination row * row stride) instead of rebuilding the generic pointer
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts4(uint32_t a, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void acc_add8_s(uint32_t addr, const float (&c)[4][4], int half) {
    const float4 v0 = lds4(addr), v1 = lds4(addr + 16);
    const float2 a = f2add(make_float2(v0.x, v0.y), make_float2(c[0][2 * half], c[0][2 * half + 1]));
    const float2 b = f2add(make_float2(v0.z, v0.w), make_float2(c[1][2 * half], c[1][2 * half + 1]));
    const float2 d = f2add(make_float2(v1.x, v1.y), make_float2(c[2][2 * half], c[2][2 * half + 1]));
    const float2 e = f2add(make_float2(v1.z, v1.w), make_float2(c[3][2 * half], c[3][2 * half + 1]));
    sts4(addr, make_float4(a.x, a.y, b.x, b.y));
    sts4(addr + 16, make_float4(d.x, d.y, e.x, e.y));
}

// row = row * scale + the lane's 8 message channels of C-fragment half `half`
__device__ __forceinline__ void acc_fma8(float* row, const float (&c)[4][4], int half, float scale) {
    float4* p = reinterpret_cast<float4*>(row);
    float4 v0 = p[0], v1 = p[1];
    v0.x = fmaf(v0.x, scale, c[0][2 * half]); v0.y = fmaf(v0.y, scale, c[0][2 * half + 1]);
    v0.z = fmaf(v0.z, scale, c[1][2 * half]); v0.w = fmaf(v0.w, scale, c[1][2 * half + 1]);
    v1.x = fmaf(v1.x, scale, c[2][2 * half]); v1.y = fmaf(v1.y, scale, c[2][2 * half + 1]);
    v1.z = fmaf(v1.z, scale, c[3][2 * half]); v1.w = fmaf(v1.w, scale, c[3][2 * half + 1]);
    p[0] = v0; p[1] = v1;
}

// lane t's 32 bytes of a split row (both k16 steps) with one 256-bit load: a whole 128-byte line per 4 lanes, so the
// L1 data stage spends one wavefront per gathered row instead of two
__device__ __forceinline__ void ld_rowh2(const uint4* __restrict__ xh, int row, int t, uint4& k0, uint4& k1) {
    const uint4* p = xh + (size_t)row * 8 + 2 * t;
#ifdef TGNN_CONV_NOALLOC
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
                 : "=r"(k0.x), "=r"(k0.y), "=r"(k0.z), "=r"(k0.w), "=r"(k1.x), "=r"(k1.y), "=r"(k1.z), "=r"(k1.w) : "l"(p));
}

// Out of line on purpose: the 3xTF32 body needs 64 registers of weight fragments; inlined, its allocation would spill
// into k_conv_h's hot loop.
template <int WN, int WARPS>
__device__ __noinline__ void conv_adj_fallback(const ConvArgs& A, float* smem) { tfx::conv_adj_body<WN, WARPS>(A, smem); }

// SPLIT = false: persistent, one warp per 64-row tile (large graphs).  SPLIT = true: one CTA per tile, its warps take
// contiguous ranges of the tile's chunks into private partial tiles that are summed in a fixed order -- the real layouts
// have ~10 tiles (N ~ 600), where one warp walking ~80 latency-bound chunks per tile would leave the GPU idle.  That geometry
// is a chain of dependent L2 accesses (measured with TGNN_ROLE_DBG: ~1350 cycles per chunk and warp, of which the MMAs and
// the scatter are ~300 each), so everything that can be requested early is: the range flags and the tile's chunk range
// together, the tile's slot indices staged in shared memory by the whole CTA in one pass (the loop then reads them at
// shared-memory latency), the root pass operands before the partial tiles are summed; and 16 warps share a tile when the
// graph has no more tiles than the GPU has SMs.
constexpr int SPLIT_CAP = 256;                     // chunks of a tile staged in shared memory (larger tiles read global memory)
constexpr int SPLIT_STAGE_BYTES = SPLIT_CAP * (CH * 4 + CH + 4);
template <int WN, int WARPS, bool SPLIT>
__global__ void __launch_bounds__(WARPS * 32, (WN == WN_BIG || WARPS > 8) ? 1 : 2)
k_conv_h(ConvArgs A) {
    constexpr int TPB = WARPS * 32;
    extern __shared__ __align__(16) float smem[];
    const long long t_start = SPLIT ? clock64() : 0ll;
    long long t_loop0 = 0, t_loop1 = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // SPLIT: S CTAs of one thread-block cluster share a tile (S = 1 without a cluster launch)
    const int S = SPLIT ? (int)cg::this_cluster().num_blocks() : 1;
    const int crank = SPLIT ? (int)cg::this_cluster().block_rank() : 0;
    const int tile_first = SPLIT ? (int)blockIdx.x / S : blockIdx.x * WARPS + warp;
    bool out_of_range;
    int c0s = 0, c1s = 0;
    if (SPLIT) {     // (requested together: a short-circuit `||` of the two flags would serialise two L2 accesses)
        const int fx = A.flag_x ? *A.flag_x : 0, fw = A.flag_w ? *A.flag_w : 0;
        if (tile_first < A.n_tiles) { c0s = __ldg(A.cptr + tile_first); c1s = __ldg(A.cptr + tile_first + 1); }
        out_of_range = (fx | fw) != 0;
    } else out_of_range = (A.flag_x && *A.flag_x) || (A.flag_w && *A.flag_w);
    if (out_of_range) {                                                       // out of the fp16 range: 3xTF32 on the fp32 rows
        conv_adj_fallback<WN, WARPS>(A, smem);                           // (same grid, same BatchNorm partial layout)
        return;
    }
    float* acc = smem + warp * (WN * XS);
    uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(acc) + 32u * (uint32_t)(threadIdx.x & 3);    // lane's 8 channels of row 0
    asm volatile("mov.u32 %0, %0;" : "+r"(acc_s));      // (opaque: kept in a register instead of being rebuilt from tid at every use)
    int* s_src = reinterpret_cast<int*>(smem + WARPS * (WN * XS));       // SPLIT: [SPLIT_CAP][16] staged slot sources ...
    int* s_type = s_src + SPLIT_CAP * CH;                                // ... [SPLIT_CAP] chunk types ...
    uint8_t* s_dst = reinterpret_cast<uint8_t*>(s_type + SPLIT_CAP);     // ... [SPLIT_CAP][16] slot destinations
    const int g = lane >> 2, t = lane & 3;
    const int gwarp = blockIdx.x * WARPS + warp, nwarp = gridDim.x * WARPS;
    double s1 = 0.0, s2 = 0.0;
    const float bias_c = __ldg(A.bias + lane);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const uint4* __restrict__ xh = A.xh;
    BFragH bf;
    int cur_type = -1;
    const int tile_step = SPLIT ? (int)gridDim.x / S : nwarp;
    const int* csrc_l = SPLIT ? nullptr : A.csrc + (lane & 15);     // persistent geometry: this lane's slot column of the chunk arrays
    const uint8_t* cdst_l = SPLIT ? nullptr : A.cdst + (lane & 15);
    if (!SPLIT) asm volatile("" : "+l"(csrc_l), "+l"(cdst_l));      // (opaque: kept in registers, not rebuilt from tid + parameters per load)

    for (int tile = tile_first; tile < A.n_tiles; tile += tile_step) {
        for (int i = lane; i < WN * XS / 4; i += 32) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        int c0, c1;
        const int* srcp = A.csrc; const uint8_t* dstp = A.cdst; const int* typep = A.ctype;
        if (SPLIT) {
            const int tt0 = tile == tile_first ? c0s : __ldg(A.cptr + tile), t1 = tile == tile_first ? c1s : __ldg(A.cptr + tile + 1);
            // this CTA's share of the tile's chunks (contiguous; the whole tile without a cluster)
            const int per_cta = (t1 - tt0 + S - 1) / S;
            const int t0 = min(tt0 + crank * per_cta, t1), t1c = min(t0 + per_cta, t1);
            const int nch = t1c - t0;
            if (nch <= SPLIT_CAP) {                                     // (CTA-uniform)
                if (tile != tile_first) __syncthreads();                // the previous tile's readers are done
                const int4* gs = reinterpret_cast<const int4*>(A.csrc + (size_t)t0 * CH);
                const uint32_t* gd = reinterpret_cast<const uint32_t*>(A.cdst + (size_t)t0 * CH);
                for (int i = threadIdx.x; i < nch * (CH / 4); i += TPB) {
                    reinterpret_cast<int4*>(s_src)[i] = __ldg(gs + i);
                    reinterpret_cast<uint32_t*>(s_dst)[i] = __ldg(gd + i);
                }
                for (int i = threadIdx.x; i < nch; i += TPB) s_type[i] = __ldg(A.ctype + t0 + i);
                __syncthreads();
                srcp = s_src - (size_t)t0 * CH; dstp = s_dst - (size_t)t0 * CH; typep = s_type - t0;
            }
            // the tile's chunks go to the CTA's warps in CONTIGUOUS ranges (chunks are sorted by type: neighbours in the list
            // mostly share their weight fragments)
            const int per = (nch + WARPS - 1) / WARPS;
            c0 = min(t0 + warp * per, t1c); c1 = min(c0 + per, t1c);
        } else { c0 = __ldg(A.cptr + tile); c1 = __ldg(A.cptr + tile + 1); }
        // (the persistent geometry reads the kernel parameters' arrays directly: no pointer registers in its hot loop)
        auto ld_src = [&](int c) -> int { return SPLIT ? srcp[(size_t)c * CH + (lane & 15)] : __ldg(csrc_l + (size_t)c * CH); };
        auto ld_dst = [&](int c) -> int { return SPLIT ? dstp[(size_t)c * CH + (lane & 15)] : __ldg(cdst_l + (size_t)c * CH); };
        auto ld_type = [&](int c) -> int { return SPLIT ? typep[c] : __ldg(A.ctype + c); };
        // software pipeline: slot indices run TWO chunks ahead of the MMAs, gathered rows ONE chunk ahead, so neither
        // the index load nor the dependent row loads are waited for in the iteration that issues them
        uint4 pre[4] = {zero4, zero4, zero4, zero4};
        int psrc = -1, pdst = 0, ptype = 0;            // chunk c
        int nsrc = -1, ndst = 0, ntype = 0;            // chunk c + 1
        if (c0 < c1) {
            psrc = ld_src(c0);
            pdst = ld_dst(c0);
            ptype = ld_type(c0);
            if (c0 + 1 < c1) {
                nsrc = ld_src(c0 + 1);
                ndst = ld_dst(c0 + 1);
                ntype = ld_type(c0 + 1);
            }
            // (an EMPTY slot gathers row 0: its product is computed and dropped -- the scatter below skips the slot -- which is
            //  cheaper than clearing eight registers and predicating the load in every iteration)
            const int sa = __shfl_sync(0xffffffffu, psrc, g), sb = __shfl_sync(0xffffffffu, psrc, g + 8);
            ld_rowh2(xh, max(sa, 0), t, pre[0], pre[1]);
            ld_rowh2(xh, max(sb, 0), t, pre[2], pre[3]);
            if (SPLIT) { load_bfrag_h(bf, A.tabH + (size_t)ptype * TG_HFRAG32, lane); cur_type = ptype; }      // with the first rows
        }
        __syncwarp();
        AQuads qa;                                         // the current chunk's rows as A-operand quads
        make_quads(pre, qa);
        if (SPLIT) t_loop0 = clock64();
        for (int c = c0; c < c1; ++c) {
            const int cdst = psrc >= 0 ? pdst : -1, type = ptype;       // this lane's slot: destination row, -1 = empty slot
            const int cn = c + 1, cn2 = c + 2;
            if (cn < c1) {                                 // rows of the next chunk (its indices arrived an iteration ago)
                const int sa = __shfl_sync(0xffffffffu, nsrc, g), sb = __shfl_sync(0xffffffffu, nsrc, g + 8);
                ld_rowh2(xh, max(sa, 0), t, pre[0], pre[1]);
                ld_rowh2(xh, max(sb, 0), t, pre[2], pre[3]);
            }
            psrc = nsrc; pdst = ndst; ptype = ntype;
            if (cn2 < c1) {                                // indices of the chunk after that
                nsrc = ld_src(cn2);
                ndst = ld_dst(cn2);
                ntype = ld_type(cn2);
            }
            if (type != cur_type) { load_bfrag_h(bf, A.tabH + (size_t)type * TG_HFRAG32, lane); cur_type = type; }
            if (ptype != type && cn < c1)          // next type's 4 KB table towards L1 (32 lines of 128 B)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(A.tabH + (size_t)ptype * TG_HFRAG32) + lane * 128));
            float m[4][4];
            chunk_mma_q(qa, bf, m);
            // rows 0..7 (group 0), then rows 8..15 (group 1): destinations are distinct inside a group
            {
                const int d = __shfl_sync(0xffffffffu, cdst, g);
                if (d >= 0) acc_add8_s(acc_s + (uint32_t)d * (XS * 4), m, 0);
            }
            __syncwarp();
            {
                const int d = __shfl_sync(0xffffffffu, cdst, g + 8);
                if (d >= 0) acc_add8_s(acc_s + (uint32_t)d * (XS * 4), m, 1);
            }
            __syncwarp();
            if (cn < c1) make_quads(pre, qa);              // take the prefetched rows over (waits for their loads here)
        }
        if (SPLIT) t_loop1 = clock64();
        const int node0 = tile * WN;
        float* tile_acc = acc;
        // SPLIT: this CTA finishes rows [row_lo, row_lo + rows_per) of the tile (all 64 without a cluster).  Root pass operands are
        // requested BEFORE the partial tiles are summed (their L2 latency hides behind the barriers); warp w owns rows
        // row_lo + 16 w .. + 15 of them
        const int rows_per = SPLIT ? WN / S : WN, row_lo = crank * rows_per;
        uint4 rcur[4] = {zero4, zero4, zero4, zero4};
        float ria = 0.f, rib = 0.f;
        const bool root_warp = SPLIT && 16 * warp < rows_per;
        const int rra = row_lo + 16 * warp + g, rrb = rra + 8;                      // tile rows of the lane's two MMA rows
        const bool rva = 16 * warp + g < rows_per, rvb = 16 * warp + g + 8 < rows_per;
        if (root_warp) {
            load_bfrag_h(bf, A.tabH + (size_t)A.n_types * TG_HFRAG32, lane); cur_type = A.n_types;      // (the chunk loop is over: bf is free)
            if (rva && node0 + rra < A.n_own) { ld_rowh2(xh, node0 + rra, t, rcur[0], rcur[1]); ria = __ldg(A.inv_deg + node0 + rra); }
            if (rvb && node0 + rrb < A.n_own) { ld_rowh2(xh, node0 + rrb, t, rcur[2], rcur[3]); rib = __ldg(A.inv_deg + node0 + rrb); }
        }
        if (SPLIT) {
            __syncthreads();
            for (int i = threadIdx.x; i < WN * XS; i += TPB) {    // fixed-order sum of the warps' partial tiles
                float v = smem[i];
#pragma unroll
                for (int w = 1; w < WARPS; ++w) v += smem[w * (WN * XS) + i];
                smem[i] = v;
            }
            tile_acc = smem;
            if (S > 1) {
                // the cluster's S partial tiles -> this CTA's rows, summed in rank order through distributed shared memory into
                // the (now free) partial-tile region of warp 1; the second cluster barrier keeps every CTA's partial tile
                // untouched until all its readers are done
                cg::cluster_group cluster = cg::this_cluster();
                cluster.sync();
                float* fin = smem + WN * XS;
                for (int i = threadIdx.x; i < rows_per * 32; i += TPB) {
                    const int off = (row_lo + (i >> 5)) * XS + (i & 31);
                    float v = 0.f;
                    for (int q = 0; q < S; ++q) v += *cluster.map_shared_rank(smem + off, q);
                    fin[off] = v;
                }
                cluster.sync();
                tile_acc = fin;
            } else __syncthreads();
        }
        // mean over in-edges and root term in ONE read-modify-write of the tile: the root pass visits every row exactly
        // once (four 16-row chunks of the tile's own rows against table entry n_types), so
        //   acc[row] = fma(acc[row], inv_deg[row], x_row @ root)
        // replaces a separate scaling pass: 128 shared-memory wavefronts less per tile (and one rounding less)
        if (SPLIT) {
            if (root_warp) {
                float m[4][4];
                chunk_mma_h(rcur, bf, m);
                if (rva) acc_fma8(tile_acc + rra * XS + 8 * t, m, 0, ria);
                if (rvb) acc_fma8(tile_acc + rrb * XS + 8 * t, m, 1, rib);
            }
        } else {
            if (cur_type != A.n_types) { load_bfrag_h(bf, A.tabH + (size_t)A.n_types * TG_HFRAG32, lane); cur_type = A.n_types; }
            for (int rc = 0; rc < WN / CH; ++rc) {
                const int na = node0 + rc * CH + g, nb = na + 8;
                uint4 cur[4] = {zero4, zero4, zero4, zero4};
                float ia = 0.f, ib = 0.f;
                if (na < A.n_own) { ld_rowh2(xh, na, t, cur[0], cur[1]); ia = __ldg(A.inv_deg + na); }
                if (nb < A.n_own) { ld_rowh2(xh, nb, t, cur[2], cur[3]); ib = __ldg(A.inv_deg + nb); }
                float m[4][4];
                chunk_mma_h(cur, bf, m);
                acc_fma8(tile_acc + (rc * CH + g) * XS + 8 * t, m, 0, ia);
                acc_fma8(tile_acc + (rc * CH + g + 8) * XS + 8 * t, m, 1, ib);
            }
        }
        if (SPLIT) __syncthreads(); else __syncwarp();
        // bias, LeakyReLU, store, statistics (lane = channel); SPLIT: the CTA's rows round-robin over its warps
        for (int r = SPLIT ? row_lo + warp : 0; r < (SPLIT ? row_lo + rows_per : WN); r += SPLIT ? WARPS : 1) {
            int node = node0 + r;
            if (node < A.n_own) {
                float v = leaky(tile_acc[r * XS + lane] + bias_c);
                if (!row_kept(A.mask, node)) v = 0.f;
                A.out[(size_t)node * F + lane] = v;
                s1 += (double)v;
                s2 += (double)v * (double)v;
            }
        }
        if (SPLIT) __syncthreads(); else __syncwarp();
    }
    if (A.part) block_part_store(A.part, s1, s2, reinterpret_cast<double*>(smem), WARPS);    // one partial row per CTA
    if (SPLIT && A.dbg && blockIdx.x == 0 && lane == 0) {                  // TGNN_ROLE_DBG=1: where the cycles of CTA 0 went
        long long* d = A.dbg + warp * 4;
        const long long t_end = clock64();
        d[0] = t_end - t_start; d[1] = t_loop0 - t_start; d[2] = t_loop1 - t_loop0; d[3] = t_end - t_loop1;
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// "X" variant: TRANSPOSED MMA roles.  In k_conv_h the gathered rows are the A operand of mma.sync.m16n8k16, whose four
// registers interleave rows g and g+8 -- two different edges, two different loads -- so the compiler assembles every A
// quad with register moves: 114 of the 282 instructions of the chunk loop (cuobjdump).  Here the per-type weights are
// the A operand (W^T, fragments straight from the table) and the gathered rows are B: a lane's B registers {k = 2t, 2t+1},
// {k = 2t+8, 2t+9} are consecutive words of ONE edge's row piece, i.e. exactly what its 256-bit load delivers -- no moves.
// The price is the accumulator layout: a lane now holds, per edge, two ADJACENT output channels (the table's row order
// makes MMA rows g / g+8 the channels 2g / 2g+1 of a 16-channel half), so the scatter into the warp's tile is 8 x
// (LDS.64, 2 FADD, STS.64) per chunk instead of 4 x (LDS.128, 4 FADD, STS.128): the same 32 wavefronts when the four edges
// a half-warp touches have destinations that differ mod 4 (row stride 40 floats; graph_build.cu arranges the slots so).
constexpr int XSX = 40;          // row stride of the accumulator tile (floats)

struct AFragX { uint4 h[2][2], l[2][2]; };       // [m-tile][k16 step]: {a0, a1, a2, a3} of W^T hi / lo
__device__ __forceinline__ void load_afrag_x(AFragX& a, const uint32_t* __restrict__ tab, int lane) {
    const uint4* p = reinterpret_cast<const uint4*>(tab) + lane;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            a.h[i][ks] = __ldg(p + ((i * 2 + ks) * 2 + 0) * 32);
            a.l[i][ks] = __ldg(p + ((i * 2 + ks) * 2 + 1) * 32);
        }
}
// rows[2 j + ks] = the lane's piece (k16 step ks) of the row of edge slot 8 j + g: {hi(c0,c1), hi(c2,c3), lo(c0,c1), lo(c2,c3)}
// m[i][j][c]: C fragment of (channel half i, edge group j): c = {edge 2t ch 2g, edge 2t+1 ch 2g, edge 2t ch 2g+1, edge 2t+1 ch 2g+1}
__device__ __forceinline__ void chunk_mma_x(const uint4 (&rows)[4], const AFragX& a, float (&m)[2][2][4]) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float sm[4] = {}, mn[2][4] = {};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint4 r = rows[2 * j + ks], wh = a.h[i][ks], wl = a.l[i][ks];
                mma_f16(sm, wh.x, wh.y, wh.z, wh.w, r.z, r.w);          // Whi . lo
                mma_f16(mn[ks], wh.x, wh.y, wh.z, wh.w, r.x, r.y);      // Whi . hi
                mma_f16(sm, wl.x, wl.y, wl.z, wl.w, r.x, r.y);          // Wlo . hi
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) m[i][j][c] = fmaf(sm[c], LO_INV, mn[0][c] + mn[1][c]);
        }
}

template <int WN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
k_conv_x(ConvArgs A) {
    extern __shared__ __align__(16) float smem[];
    if ((A.flag_x && *A.flag_x) || (A.flag_w && *A.flag_w)) {            // out of the fp16 range: 3xTF32 on the fp32 rows
        conv_adj_fallback<WN, WARPS>(A, smem);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* acc = smem + warp * (WN * XSX);
    const int g = lane >> 2, t = lane & 3;
    const int gwarp = blockIdx.x * WARPS + warp, nwarp = gridDim.x * WARPS;
    double s1 = 0.0, s2 = 0.0;
    const float bias_c = __ldg(A.bias + lane);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const uint4* __restrict__ xh = A.xh;
    AFragX af;
    int cur_type = -1;
    for (int tile = gwarp; tile < A.n_tiles; tile += nwarp) {
        for (int i = lane; i < WN * XSX / 4; i += 32) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c0 = __ldg(A.cptr + tile), c1 = __ldg(A.cptr + tile + 1);
        // software pipeline as in k_conv_h: slot indices TWO chunks ahead, gathered rows ONE chunk ahead
        uint4 pre[4] = {zero4, zero4, zero4, zero4};
        int psrc = -1, pdst = 0, ptype = 0, nsrc = -1, ndst = 0, ntype = 0;
        if (c0 < c1) {
            psrc = __ldg(A.csrc + (size_t)c0 * CH + (lane & 15));
            pdst = __ldg(A.cdst + (size_t)c0 * CH + (lane & 15));
            ptype = __ldg(A.ctype + c0);
            if (c0 + 1 < c1) {
                nsrc = __ldg(A.csrc + (size_t)(c0 + 1) * CH + (lane & 15));
                ndst = __ldg(A.cdst + (size_t)(c0 + 1) * CH + (lane & 15));
                ntype = __ldg(A.ctype + c0 + 1);
            }
            const int sa = __shfl_sync(0xffffffffu, psrc, g), sb = __shfl_sync(0xffffffffu, psrc, g + 8);
            if (sa >= 0) ld_rowh2(xh, sa, t, pre[0], pre[1]);
            if (sb >= 0) ld_rowh2(xh, sb, t, pre[2], pre[3]);
        }
        __syncwarp();
        for (int c = c0; c < c1; ++c) {
            const uint4 cur[4] = {pre[0], pre[1], pre[2], pre[3]};
            const int cdst = psrc >= 0 ? pdst : -1, type = ptype;       // destination row of this lane's slot, -1 = empty slot
            if (c + 1 < c1) {
                const int sa = __shfl_sync(0xffffffffu, nsrc, g), sb = __shfl_sync(0xffffffffu, nsrc, g + 8);
                pre[0] = pre[1] = pre[2] = pre[3] = zero4;
                if (sa >= 0) ld_rowh2(xh, sa, t, pre[0], pre[1]);
                if (sb >= 0) ld_rowh2(xh, sb, t, pre[2], pre[3]);
            }
            psrc = nsrc; pdst = ndst; ptype = ntype;
            if (c + 2 < c1) {
                nsrc = __ldg(A.csrc + (size_t)(c + 2) * CH + (lane & 15));
                ndst = __ldg(A.cdst + (size_t)(c + 2) * CH + (lane & 15));
                ntype = __ldg(A.ctype + c + 2);
            }
            if (type != cur_type) { load_afrag_x(af, A.tabX + (size_t)type * TG_HFRAG32, lane); cur_type = type; }
            if (ptype != type && c + 1 < c1)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(A.tabX + (size_t)ptype * TG_HFRAG32) + lane * 128));
            float m[2][2][4];
            chunk_mma_x(cur, af, m);
            // scatter: slot 8 j + 2 t + u; destinations are distinct inside a group of 8 slots (j), the groups are separated by
            // a warp barrier
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int d = __shfl_sync(0xffffffffu, cdst, 8 * j + 2 * t + u);
                    if (d >= 0) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float2* p = reinterpret_cast<float2*>(acc + d * XSX + 16 * i + 2 * g);
                            float2 v = *p;
                            v.x += m[i][j][u]; v.y += m[i][j][2 + u];
                            *p = v;
                        }
                    }
                }
                __syncwarp();
            }
        }
        const int node0 = tile * WN;
        // mean over in-edges and root term in one read-modify-write of the tile (as k_conv_h): acc = fma(acc, inv_deg, x_row @ root)
        if (cur_type != A.n_types) { load_afrag_x(af, A.tabX + (size_t)A.n_types * TG_HFRAG32, lane); cur_type = A.n_types; }
        for (int rc = 0; rc < WN / CH; ++rc) {
            const int na = node0 + rc * CH + g, nb = na + 8;
            uint4 cur[4] = {zero4, zero4, zero4, zero4};
            if (na < A.n_own) ld_rowh2(xh, na, t, cur[0], cur[1]);
            if (nb < A.n_own) ld_rowh2(xh, nb, t, cur[2], cur[3]);
            float m[2][2][4];
            chunk_mma_x(cur, af, m);
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int r = rc * CH + 8 * j + 2 * t + u;
                    const float id = node0 + r < A.n_own ? __ldg(A.inv_deg + node0 + r) : 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float2* p = reinterpret_cast<float2*>(acc + r * XSX + 16 * i + 2 * g);
                        float2 v = *p;
                        v.x = fmaf(v.x, id, m[i][j][u]); v.y = fmaf(v.y, id, m[i][j][2 + u]);
                        *p = v;
                    }
                }
        }
        __syncwarp();
        // bias, LeakyReLU, store, statistics (lane = channel)
        for (int r = 0; r < WN; ++r) {
            const int node = node0 + r;
            if (node < A.n_own) {
                float v = leaky(acc[r * XSX + lane] + bias_c);
                if (!row_kept(A.mask, node)) v = 0.f;
                A.out[(size_t)node * F + lane] = v;
                s1 += (double)v;
                s2 += (double)v * (double)v;
            }
        }
        __syncwarp();
    }
    if (A.part) block_part_store(A.part, s1, s2, reinterpret_cast<double*>(smem), WARPS);    // one partial row per CTA
}

}  // namespace

void launch_conv_h(const ConvArgs& a, int sm_count, cudaStream_t st) {
    static PerDeviceOnce once;
    const size_t smem_small = (size_t)8 * WN_SMALL * XS * sizeof(float), smem_big = (size_t)12 * WN_BIG * XS * sizeof(float);
    const size_t smem_split8 = smem_small + SPLIT_STAGE_BYTES, smem_split16 = 2 * smem_small + SPLIT_STAGE_BYTES;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_h<WN_SMALL, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small));
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_h<WN_SMALL, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_split8));
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_h<WN_SMALL, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_split16));
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_h<WN_BIG, 12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
    });
    // same grid as k_conv_adj: the BatchNorm partial layout is shared by the two kernels
    const ConvGeom g = conv_geom(a.n_tiles, a.wn, sm_count, a.n_chunks);
    const bool w16_off = getenv("TGNN_CONV_W16") && std::string(getenv("TGNN_CONV_W16")) == "0";
    if (a.wn == WN_BIG) k_conv_h<WN_BIG, 12, false><<<g.blocks, 12 * 32, smem_big, st>>>(a);
    else if (g.split && g.cluster > 1) {                       // fewer tiles than SMs: a cluster of CTAs per tile
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(g.blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem_split8; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = g.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        TGNN_CUDA(cudaLaunchKernelEx(&cfg, k_conv_h<WN_SMALL, 8, true>, a));
    }
    else if (g.split && g.blocks <= sm_count && !w16_off) k_conv_h<WN_SMALL, 16, true><<<g.blocks, 512, smem_split16, st>>>(a);   // a tile per SM: 16 warps share it
    else if (g.split) k_conv_h<WN_SMALL, 8, true><<<g.blocks, 256, smem_split8, st>>>(a);
    else k_conv_h<WN_SMALL, 8, false><<<g.blocks, 256, smem_small, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

// transposed-roles variant: large graphs only (the same persistent grid as k_conv_h's non-split geometry)
void launch_conv_x(const ConvArgs& a, int sm_count, cudaStream_t st) {
    static PerDeviceOnce once;
    const size_t smem = (size_t)8 * WN_SMALL * XSX * sizeof(float);
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_x<WN_SMALL, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
    const ConvGeom g = conv_geom(a.n_tiles, a.wn, sm_count, a.n_chunks);
    TGNN_CHECK(a.wn == WN_SMALL && !g.split, "internal: k_conv_x needs 64-row tiles and the persistent geometry");
    k_conv_x<WN_SMALL, 8><<<g.blocks, 256, smem, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
