// Adjacency branch, "Z" kernel: destination-tile passes on the 5th-generation tensor cores with the A operand in
// TENSOR MEMORY and the neighbour rows staged once per tile in a shared-memory WINDOW
// (graph_networks/layers/edge_conv.py:24-27 of the reference; PyG NNConv(aggr="mean") + root + bias, LeakyReLU,
// BatchNorm partial sums -- same arithmetic contract as k_conv_h / k_conv_s).
//
// Formulation (the S format of graph_build.cu): for a tile of 128 destinations and one edge type t,
//   Z_t[r] = sum of the source rows of destination r's type-t in-edges   (zero row if none)
//   D[128 x 32] += Z_t W_t                                              (one PASS per type present in the tile)
// so the accumulator lives in TMEM over all passes of the tile: no scatter of per-edge messages, no read-modify-write
// of a shared-memory tile (what bounds k_conv_h: 32 + 16 + 13 L1 wavefronts per 16 edges and an L2-latency gather).
//
// What is new against k_conv_s (which built Z_t in shared memory from a cp.async ring of rows gathered out of L2):
//   * WINDOW: the distinct source rows of a tile's in-edges (lattice, deg 32: 4096 edges -> 938 rows in 7 runs) are
//     brought into shared memory once per tile by TMA bulk copies (cp.async.bulk -> UBLKCP, one per contiguous run,
//     completion on an mbarrier); every edge is a uint16 window offset.  L2 -> SM row traffic / 4.4.
//   * A IN TMEM (tcgen05.mma with the A operand in tensor memory): a gather thread owns ONE destination row of the
//     pass: it reads its source row(s) from the window with 8 conflict-free LDS.128 (lane l starts at 16-byte chunk
//     l mod 8, un-rotated in registers), sums multi-edges in fp32, splits hi = top 19 bits / lo = x - hi and writes
//     the row straight into the A stage with tcgen05.st.  The tensor core never reads A from shared memory, so the
//     shared-memory pipe carries 1 wavefront per edge + the 8 KB weight image per pass and nothing else.
//   * 3xTF32 on fp32 rows (hi.Whi + lo.Whi + hi.Wlo, 12 MMAs of M=128, N=32, K=8 per pass): full fp32 range, no
//     fp16 range flags, no split copy of b1 needed.
// Warp roles (16 warps, one persistent CTA per SM):
//   warp 0      window producer: per tile <= 32 bulk copies (runs of rows)
//   warp 1      pass producer: per pass the pre-swizzled weight image (8 KB), the 129-entry row-offset table and the
//               pass's window offsets, into a 4-slot ring
//   warp 2      MMA issuer (one lane)
//   warps 4-11  gather: two groups of 4 warps (TMEM lane quarter = warp % 4) take alternate passes; 6 A stages in TMEM
//   warps 12-15 epilogue: tcgen05.ld, * 1/deg + root + bias, LeakyReLU, store, BatchNorm partial sums (fp64); the
//               accumulators are double buffered so the epilogue of tile i overlaps the passes of tile i+1
// The root term x_i root is one more pass (A = the tile's own rows, which the window also holds) into a second
// accumulator.  Graphs whose tiles are not local enough for a window (or tiny graphs) keep k_conv_h.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "layouts.cuh"
#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {
using namespace tc;

constexpr int Z_NG = 2;                          // gather groups of 4 warps
constexpr int Z_NS = 4;                          // pass slots (weight image + offset tables) in flight
constexpr int Z_NSTA = 6;                        // A operand stages in TMEM (64 columns each: hi | lo)
constexpr int W_PRODW = 0, W_PRODP = 1, W_MMA = 2, W_G0 = 4, W_EPI0 = W_G0 + 4 * Z_NG;
constexpr int CZ_THREADS = (W_EPI0 + 4) * 32;
constexpr int SB_TILE = 4096;                    // bytes of one weight image (hi or lo): 32 rows x 128 B
constexpr int SLOT_OFF = 2 * SB_TILE;            // row-offset table of the pass (S_OFF_STRIDE uint16)
constexpr int SLOT_LOC = SLOT_OFF + 288;         // window offsets of the pass's edges (16-byte aligned superset)
constexpr int SLOT_BYTES = 9216;                 // 8192 + 288 + (ZW_MAX_PASS + 16) * 2 <= 9216, multiple of 1024
static_assert(SLOT_LOC + (ZW_MAX_PASS + 16) * 2 <= SLOT_BYTES, "pass slot too small");
constexpr int OFF_SLOTS = 0;
constexpr int OFF_WIN = OFF_SLOTS + Z_NS * SLOT_BYTES;          // [ZW_WMAX][128 B]
constexpr int OFF_EPI = OFF_WIN + ZW_WMAX * 128;                // scratch [4][32*33] float, red [4][2][32] double
constexpr int CZ_SMEM = OFF_EPI + 4 * 32 * 33 * 4 + 4 * 2 * 32 * 8 + 1024;
constexpr int TM_A0 = 128;                       // TMEM columns: [0,128) two accumulator buffers {agg 32 | root 32}, then the A stages
static_assert(TM_A0 + Z_NSTA * 64 <= 512, "TMEM columns");

struct ConvZArgs {
    const float* xin;            // [n_rows][32]
    const float* tabS;           // [K+1][hi|lo] swizzled images of W_t^T [32 n][32 k]  (entry K = root^T)
    int n_types;
    const int* pptr; const int* ptype; const int* pbase; const unsigned short* off;
    const int* zmeta; const int* zseg; const unsigned short* zloc;
    const float* inv_deg; const float* bias;
    float* out; double* part; int* error_flag;
    const uint8_t* mask;
    long long* dbg;              // optional per-warp {cycles, wait0, wait1, wait2} of CTA 0 (TGNN_ROLE_DBG=1)
    int n_own, n_tiles;
};

// tcgen05.mma with A in tensor memory (lane = row, 32-bit column = K element), B from a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
}
__device__ __forceinline__ float4 sel4(bool p, const float4& a, const float4& b) {
    return make_float4(p ? a.x : b.x, p ? a.y : b.y, p ? a.z : b.z, p ? a.w : b.w);
}

__global__ void __launch_bounds__(CZ_THREADS, 1)
k_conv_z(ConvZArgs A) {
    long long w0 = 0, w1 = 0, w2 = 0;                  // cycles spent in this role's barrier waits
    const long long t_start = clock64();
    extern __shared__ uint8_t smem_raw[];
    // win_full, win_empty, slot_full[NS], slot_empty[NS], a_full[NSTA], a_empty[NSTA], acc_full[2], acc_empty[2]
    __shared__ __align__(8) uint64_t bars[2 + 2 * Z_NS + 2 * Z_NSTA + 4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int timeout_flag;
    __shared__ int slot_shift[Z_NS];                   // first edge of the pass inside the slot's (aligned) offset copy
    __shared__ int win_self;                           // window row of the tile's first own row
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_wf = smem_u32(&bars[0]), bar_we = smem_u32(&bars[1]);
    const uint32_t bar_sf = smem_u32(&bars[2]), bar_se = smem_u32(&bars[2 + Z_NS]);
    const uint32_t bar_af = smem_u32(&bars[2 + 2 * Z_NS]), bar_ae = smem_u32(&bars[2 + 2 * Z_NS + Z_NSTA]);
    const uint32_t bar_cf = smem_u32(&bars[2 + 2 * Z_NS + 2 * Z_NSTA]), bar_ce = bar_cf + 16;

    if (tid == 0) {
        mbar_init(bar_wf, 1); mbar_init(bar_we, 4 * Z_NG);
        for (int i = 0; i < Z_NS; ++i) { mbar_init(bar_sf + 8 * i, 1); mbar_init(bar_se + 8 * i, 4 + 1); }      // 4 gather warps + the MMA commit
        for (int i = 0; i < Z_NSTA; ++i) { mbar_init(bar_af + 8 * i, 4); mbar_init(bar_ae + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_cf + 8 * i, 1); mbar_init(bar_ce + 8 * i, 4); }
        timeout_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == W_PRODW) {
        // ===================== window producer =====================
        int it = 0;
        for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++it) {
            const int4 m = __ldg(reinterpret_cast<const int4*>(A.zmeta) + tile);          // {runs, window rows, window row of the first own row, edges}
            int2 sg = make_int2(0, 0);
            int next = m.y;
            if (lane < m.x) {
                sg = __ldg(reinterpret_cast<const int2*>(A.zseg) + (size_t)tile * ZW_MAXSEG + lane);
                if (lane + 1 < m.x) next = __ldg(A.zseg + ((size_t)tile * ZW_MAXSEG + lane + 1) * 2 + 1);
            }
            if (!TGNN_TIMED(w0, mbar_wait_relaxed(bar_we, (uint32_t)((it & 1) ^ 1)))) { timeout_flag = 1; break; }
            if (lane == 0) {
                win_self = m.z;
                mbar_arrive_expect_tx(bar_wf, (uint32_t)m.y * 128u);      // (release: win_self is visible to the waiters)
            }
            __syncwarp();
            if (lane < m.x)
                bulk_g2s(sbase + OFF_WIN + (uint32_t)sg.y * 128u, A.xin + (size_t)sg.x * F, (uint32_t)(next - sg.y) * 128u, bar_wf);
        }
    } else if (warp == W_PRODP) {
        // ===================== pass producer =====================
        int s = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x) {
            const int p0 = __ldg(A.pptr + tile), np = __ldg(A.pptr + tile + 1) - p0;
            for (int qb = 0; qb <= np && ok; qb += 32) {
                // lane l looks up pass qb + l; lane 0 issues the copies in pass order
                const int q = qb + lane;
                int type = A.n_types, eb = 0, ee = 0;
                if (q < np) { type = __ldg(A.ptype + p0 + q); eb = __ldg(A.pbase + p0 + q); ee = __ldg(A.pbase + p0 + q + 1); }
                const int nq = min(32, np + 1 - qb);
                for (int j = 0; j < nq; ++j, ++s) {
                    const int ty = __shfl_sync(0xffffffffu, type, j), b0 = __shfl_sync(0xffffffffu, eb, j), b1 = __shfl_sync(0xffffffffu, ee, j);
                    const bool root = qb + j == np;
                    const int slot = s % Z_NS;
                    if (!TGNN_TIMED(w0, mbar_wait_relaxed(bar_se + 8 * slot, (uint32_t)(((s / Z_NS) & 1) ^ 1)))) { ok = false; break; }
                    if (lane == 0) {
                        const uint32_t dst = sbase + OFF_SLOTS + (uint32_t)slot * SLOT_BYTES, bar = bar_sf + 8 * slot;
                        const int a0 = b0 & ~7, a1 = (b1 + 7) & ~7;
                        slot_shift[slot] = b0 - a0;
                        const uint32_t loc_bytes = root ? 0u : (uint32_t)(a1 - a0) * 2u;
                        mbar_arrive_expect_tx(bar, 2u * SB_TILE + (root ? 0u : (uint32_t)S_OFF_STRIDE * 2u) + loc_bytes);
                        bulk_g2s(dst, A.tabS + (size_t)ty * 2048, 2u * SB_TILE, bar);
                        if (!root) {
                            bulk_g2s(dst + SLOT_OFF, A.off + (size_t)(p0 + qb + j) * S_OFF_STRIDE, (uint32_t)S_OFF_STRIDE * 2u, bar);
                            if (loc_bytes) bulk_g2s(dst + SLOT_LOC, A.zloc + a0, loc_bytes, bar);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (!ok) timeout_flag = 1;
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t IDESC = umma_idesc_tf32(32);
            const uint64_t dB = umma_desc_sw128(sbase + OFF_SLOTS);
            int s = 0, it = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x, ++it) {
                const int np = __ldg(A.pptr + tile + 1) - __ldg(A.pptr + tile);
                const int ab = it & 1;
                if (!TGNN_TIMED(w2, mbar_wait(bar_ce + 8 * ab, (uint32_t)(((it >> 1) & 1) ^ 1)))) { ok = false; break; }
                for (int q = 0; q <= np; ++q, ++s) {
                    const bool root = q == np;
                    const int slot = s % Z_NS, stg = s % Z_NSTA;
                    if (!TGNN_TIMED(w0, mbar_wait(bar_sf + 8 * slot, (uint32_t)((s / Z_NS) & 1)))) { ok = false; break; }
                    if (!TGNN_TIMED(w1, mbar_wait(bar_af + 8 * stg, (uint32_t)((s / Z_NSTA) & 1)))) { ok = false; break; }
                    fence_proxy_async();
                    tc_fence_after();
                    const uint64_t dbh = dB + (uint64_t)(slot * (SLOT_BYTES / 16)), dbl = dbh + SB_TILE / 16;
                    const uint32_t a_hi = tmem_base + (uint32_t)(TM_A0 + stg * 64), a_lo = a_hi + 32u;
                    const uint32_t tmem_d = tmem_base + (uint32_t)(ab * 64) + (root ? 32u : 0u);
                    const uint32_t first = (root || q == 0) ? 0u : 1u;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {       // 8 tf32 = 8 TMEM columns of A = 32 bytes = 2 descriptor units of B inside the swizzle atom
                        umma_tf32_ts(tmem_d, a_lo + 8 * ks, dbh + 2 * ks, IDESC, ks == 0 ? first : 1u);
                        umma_tf32_ts(tmem_d, a_hi + 8 * ks, dbl + 2 * ks, IDESC, 1u);
                        umma_tf32_ts(tmem_d, a_hi + 8 * ks, dbh + 2 * ks, IDESC, 1u);
                    }
                    umma_commit(bar_se + 8 * slot);        // weight image free
                    umma_commit(bar_ae + 8 * stg);         // A stage free
                    if (root) umma_commit(bar_cf + 8 * ab);
                }
            }
            if (!ok) timeout_flag = 1;
        }
    } else if (warp >= W_G0 && warp < W_EPI0) {
        // ===================== gather: thread = destination row of the pass =====================
        const int grp = (warp - W_G0) >> 2, q4 = warp & 3;
        const int r = 32 * q4 + lane, rho = lane & 7;
        const uint32_t t_lane = (uint32_t)(32 * q4) << 16;
        const uint32_t win = sbase + OFF_WIN;
        uint32_t co[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) co[k] = (uint32_t)(((k + rho) & 7) << 4);
        int s = 0, it = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x, ++it) {
            const int np = __ldg(A.pptr + tile + 1) - __ldg(A.pptr + tile);
            const bool live = tile * S_BM + r < A.n_own;
            if (!TGNN_TIMED(w2, mbar_wait(bar_wf, (uint32_t)(it & 1)))) { ok = false; break; }
            const int self_loc = win_self;
            for (int q = 0; q <= np; ++q, ++s) {
                if (s % Z_NG != grp) continue;
                const bool root = q == np;
                const int slot = s % Z_NS, stg = s % Z_NSTA;
                if (!TGNN_TIMED(w0, mbar_wait(bar_sf + 8 * slot, (uint32_t)((s / Z_NS) & 1)))) { ok = false; break; }
                const uint32_t sl = sbase + OFF_SLOTS + (uint32_t)slot * SLOT_BYTES;
                int cnt;
                uint32_t lp = 0;
                if (!root) {
                    const int e0 = (int)lds_u16(sl + SLOT_OFF + 2 * r), e1 = (int)lds_u16(sl + SLOT_OFF + 2 * r + 2);
                    cnt = e1 - e0;
                    lp = sl + SLOT_LOC + 2u * (uint32_t)(slot_shift[slot] + e0);
                } else cnt = live ? 1 : 0;
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cnt > 0) {
                    const uint32_t rowa = win + (root ? (uint32_t)(self_loc + r) : lds_u16(lp)) * 128u;
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = lds128f(rowa + co[k]);
                }
                for (int t = 1; t < cnt; ++t) {           // multi-edges of one (row, type): fp32 sum, like the reference's scatter
                    const uint32_t rowa = win + lds_u16(lp + 2u * (uint32_t)t) * 128u;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float4 w = lds128f(rowa + co[k]);
                        v[k].x += w.x; v[k].y += w.y; v[k].z += w.z; v[k].w += w.w;
                    }
                }
                // v[k] holds chunk (k + rho) & 7 of the row: rotate back by rho in three select stages
                float4 a[8], b[8], u[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] = sel4(rho & 1, v[(c + 7) & 7], v[c]);
#pragma unroll
                for (int c = 0; c < 8; ++c) b[c] = sel4(rho & 2, a[(c + 6) & 7], a[c]);
#pragma unroll
                for (int c = 0; c < 8; ++c) u[c] = sel4(rho & 4, b[(c + 4) & 7], b[c]);
                // the A stage was last read by the MMAs of pass s - NSTA
                if (!TGNN_TIMED(w1, mbar_wait(bar_ae + 8 * stg, (uint32_t)(((s / Z_NSTA) & 1) ^ 1)))) { ok = false; break; }
                tc_fence_after();
                const uint32_t ta = tmem_base + t_lane + (uint32_t)(TM_A0 + stg * 64);
                uint32_t hi[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    hi[4 * c + 0] = __float_as_uint(u[c].x) & 0xFFFFE000u; hi[4 * c + 1] = __float_as_uint(u[c].y) & 0xFFFFE000u;
                    hi[4 * c + 2] = __float_as_uint(u[c].z) & 0xFFFFE000u; hi[4 * c + 3] = __float_as_uint(u[c].w) & 0xFFFFE000u;
                }
                tmem_st32(ta, hi);
#pragma unroll
                for (int c = 0; c < 8; ++c) {              // lo = x - hi is exact; the tensor core truncates it to TF32 itself
                    hi[4 * c + 0] = __float_as_uint(u[c].x - __uint_as_float(hi[4 * c + 0]));
                    hi[4 * c + 1] = __float_as_uint(u[c].y - __uint_as_float(hi[4 * c + 1]));
                    hi[4 * c + 2] = __float_as_uint(u[c].z - __uint_as_float(hi[4 * c + 2]));
                    hi[4 * c + 3] = __float_as_uint(u[c].w - __uint_as_float(hi[4 * c + 3]));
                }
                tmem_st32(ta + 32u, hi);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_af + 8 * stg); mbar_arrive(bar_se + 8 * slot); }
            }
            if (!ok) break;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_we);                    // this warp is done with the tile's window
        }
        if (!ok) timeout_flag = 1;
    } else if (warp >= W_EPI0) {
        // ===================== epilogue warps: TMEM lane quarter q4 = warp % 4 =====================
        const int q4 = warp & 3, etid = (warp - W_EPI0) * 32 + lane;
        const uint32_t sc = sbase + OFF_EPI + (uint32_t)q4 * (32 * 33 * 4);
        const uint32_t red = sbase + OFF_EPI + 4 * 32 * 33 * 4;
        int it = 0;
        for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const int np = __ldg(A.pptr + tile + 1) - __ldg(A.pptr + tile);
            if (!TGNN_TIMED(w0, mbar_wait_relaxed(bar_cf + 8 * ab, (uint32_t)((it >> 1) & 1)))) { timeout_flag = 1; break; }
            tc_fence_after();
            const int row = tile * S_BM + 32 * q4 + lane;
            const bool live = row < A.n_own;
            uint32_t vt[32], vr[32];
            const uint32_t tbase = tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)(ab * 64);
            tmem_ld32(tbase + 32u, vr);
            if (np > 0) tmem_ld32(tbase, vt);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ce + 8 * ab);          // accumulator buffer free for the tile after next
            const float idg = live ? __ldg(A.inv_deg + row) : 0.f;
            const bool kept = live && row_kept(A.mask, row);
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float t = np > 0 ? __uint_as_float(vt[j]) : 0.f;
                o[j] = kept ? leaky(fmaf(t, idg, __uint_as_float(vr[j])) + __ldg(A.bias + j)) : 0.f;
            }
            if (live) {
                float4* dst = reinterpret_cast<float4*>(A.out + (size_t)row * F);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            if (A.part) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sts_f32(sc + 4 * (lane * 33 + j), o[j]);
                __syncwarp();
                int nv = A.n_own - (tile * S_BM + 32 * q4);
                nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
                double s1 = 0.0, s2 = 0.0;
                for (int rr = 0; rr < nv; ++rr) { const double x = (double)lds_f32(sc + 4 * (rr * 33 + lane)); s1 += x; s2 += x * x; }
                sts_f64(red + 8 * ((q4 * 2 + 0) * 32 + lane), s1);
                sts_f64(red + 8 * ((q4 * 2 + 1) * 32 + lane), s2);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (etid < 64) {
                    const int qq = etid >> 5, cc = etid & 31;
                    A.part[(size_t)tile * 64 + etid] = ((lds_f64(red + 8 * ((0 * 2 + qq) * 32 + cc)) + lds_f64(red + 8 * ((1 * 2 + qq) * 32 + cc))) +
                                                        lds_f64(red + 8 * ((2 * 2 + qq) * 32 + cc))) + lds_f64(red + 8 * ((3 * 2 + qq) * 32 + cc));
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    if (A.dbg && blockIdx.x == 0 && lane == 0) {
        long long* d = A.dbg + warp * 4;
        d[0] = clock64() - t_start; d[1] = w0; d[2] = w1; d[3] = w2;
    }
    tc_fence_before();
    __syncthreads();
    if (timeout_flag && tid == 0) { *reinterpret_cast<volatile int*>(A.error_flag) = TGNN_DEVERR_PIPELINE; __threadfence_system(); }   // mapped host word
    if (warp == W_MMA) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- window builder: one CTA per 128-row tile sorts the sources of the tile's in-edges (S-format order) together with
// ---- its own rows (the root pass reads them from the window too), cuts them into contiguous runs (gaps of <= ZW_GAP
// ---- rows are loaded rather than split) and writes every edge's window offset -------------------------------------------
constexpr int ZW_ITEMS = ZW_CAP / 256;

__global__ void __launch_bounds__(256)
k_zw_build(const int* __restrict__ pptr, const int* __restrict__ pbase, const int* __restrict__ s_src, int n_own,
           int* __restrict__ meta, int* __restrict__ seg, unsigned short* __restrict__ loc, int* __restrict__ n_bad) {
    using Sort = cub::BlockRadixSort<int, 256, ZW_ITEMS>;
    using Scan = cub::BlockScan<int, 256>;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ int srt[ZW_CAP];
    __shared__ int seg_first[ZW_MAXSEG], seg_lbase[ZW_MAXSEG];
    __shared__ int s_nseg, s_rows;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int node0 = tile * S_BM, node1 = min(node0 + S_BM, n_own);
    const int e0 = pbase[pptr[tile]], e1 = pbase[pptr[tile + 1]];
    const int ne = e1 - e0, n_items = ne + (node1 - node0);
    int* m = meta + 4 * tile;
    if (n_items > ZW_CAP) {
        if (tid == 0) { m[0] = 0; m[1] = 0; m[2] = 0; m[3] = ne; atomicAdd(n_bad, 1); }
        return;
    }
    int items[ZW_ITEMS];
#pragma unroll
    for (int k = 0; k < ZW_ITEMS; ++k) {
        const int i = tid * ZW_ITEMS + k;
        items[k] = i < ne ? s_src[e0 + i] : (i < n_items ? node0 + (i - ne) : 0x7fffffff);
    }
    Sort(tmp.sort).Sort(items);
#pragma unroll
    for (int k = 0; k < ZW_ITEMS; ++k) srt[tid * ZW_ITEMS + k] = items[k];
    __syncthreads();
    // run starts and the rows skipped in front of each run ("jump"): local(row) = row - (sum of jumps up to it)
    int jump[ZW_ITEMS], start[ZW_ITEMS], jsum = 0, ssum = 0;
#pragma unroll
    for (int k = 0; k < ZW_ITEMS; ++k) {
        const int i = tid * ZW_ITEMS + k;
        const int v = items[k], prev = i > 0 ? srt[i - 1] : 0;
        const bool valid = i < n_items;
        const bool st = valid && (i == 0 || v - prev - 1 > ZW_GAP);
        start[k] = st ? 1 : 0;
        jump[k] = st ? (i == 0 ? v : v - prev - 1) : 0;
        jsum += jump[k]; ssum += start[k];
    }
    int jpre, spre;
    Scan(tmp.scan).ExclusiveSum(jsum, jpre);
    __syncthreads();
    Scan(tmp.scan).ExclusiveSum(ssum, spre);
    if (tid == 0) { s_nseg = 0; s_rows = 0; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ZW_ITEMS; ++k) {
        const int i = tid * ZW_ITEMS + k;
        jpre += jump[k]; spre += start[k];
        if (start[k] && spre <= ZW_MAXSEG) { seg_first[spre - 1] = items[k]; seg_lbase[spre - 1] = items[k] - jpre; }
        if (i == n_items - 1) { s_nseg = spre; s_rows = items[k] - jpre + 1; }
    }
    __syncthreads();
    const int nseg = s_nseg, rows = s_rows;
    if (nseg > ZW_MAXSEG || rows > ZW_WMAX) {
        if (tid == 0) { m[0] = 0; m[1] = 0; m[2] = 0; m[3] = ne; atomicAdd(n_bad, 1); }
        return;
    }
    auto local_of = [&](int row) {
        int lo = 0, hi = nseg - 1;                       // last run whose first row is <= row
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (seg_first[mid] <= row) lo = mid; else hi = mid - 1; }
        return seg_lbase[lo] + (row - seg_first[lo]);
    };
    for (int i = tid; i < ne; i += 256) loc[e0 + i] = (unsigned short)local_of(s_src[e0 + i]);
    if (tid < nseg) { seg[((size_t)tile * ZW_MAXSEG + tid) * 2] = seg_first[tid]; seg[((size_t)tile * ZW_MAXSEG + tid) * 2 + 1] = seg_lbase[tid]; }
    if (tid == 0) { m[0] = nseg; m[1] = rows; m[2] = local_of(node0); m[3] = ne; }
}

}  // namespace

// Builds the windows of the S format already in g.  Returns the number of tiles WITHOUT a window (the Z kernel needs 0).
int build_z_windows(Graph& g, Scratch& sc, cudaStream_t st) {
    g.zw_meta.reserve((size_t)g.s_tiles * 4 * sizeof(int));
    g.zw_seg.reserve((size_t)g.s_tiles * ZW_MAXSEG * 2 * sizeof(int));
    g.z_loc.reserve(((size_t)g.e_adj + 64) * sizeof(unsigned short));
    int* nb = sc.get<int>(1);
    TGNN_CUDA(cudaMemsetAsync(nb, 0, sizeof(int), st));
    k_zw_build<<<g.s_tiles, 256, 0, st>>>(g.s_pptr.as<int>(), g.s_pbase.as<int>(), g.s_src.as<int>(), (int)g.n_own,
                                          g.zw_meta.as<int>(), g.zw_seg.as<int>(), g.z_loc.as<unsigned short>(), nb);
    TGNN_CUDA(cudaGetLastError());
    int n_bad = 0;
    TGNN_CUDA(cudaMemcpyAsync(&n_bad, nb, sizeof(int), cudaMemcpyDeviceToHost, st));
    TGNN_CUDA(cudaStreamSynchronize(st));
    return n_bad;
}

int conv_z_blocks(int s_tiles, int sm_count) { return s_tiles < sm_count ? (s_tiles < 1 ? 1 : s_tiles) : sm_count; }

void launch_conv_z(const ConvArgs& c, const Graph& g, const float* tabS, int* error_flag, int sm_count, cudaStream_t st, long long* dbg) {
    static PerDeviceOnce once;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_z, cudaFuncAttributeMaxDynamicSharedMemorySize, CZ_SMEM));
    });
    TGNN_CHECK(g.has_z, "internal: k_conv_z without its windows");
    ConvZArgs a{};
    a.xin = c.xin; a.tabS = tabS; a.n_types = g.n_types;
    a.pptr = g.s_pptr.as<int>(); a.ptype = g.s_ptype.as<int>(); a.pbase = g.s_pbase.as<int>();
    a.off = g.s_off.as<unsigned short>();
    a.zmeta = g.zw_meta.as<int>(); a.zseg = g.zw_seg.as<int>(); a.zloc = g.z_loc.as<unsigned short>();
    a.inv_deg = c.inv_deg; a.bias = c.bias; a.out = c.out; a.part = c.part; a.error_flag = error_flag; a.mask = c.mask;
    a.n_own = c.n_own; a.n_tiles = g.s_tiles; a.dbg = dbg;
    k_conv_z<<<conv_z_blocks(g.s_tiles, sm_count), CZ_THREADS, CZ_SMEM, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
