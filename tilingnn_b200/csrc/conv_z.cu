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
//   * Two arithmetic variants of the same pipeline.  A tcgen05.mma costs ~44.6 cycles whatever its N <= 64 and kind
//     (scripts/micro/umma_ts_rate.cu, measured on B200), so the instruction COUNT per pass is what matters:
//       k_conv_z<true>  (default): the window holds rows of the fp16-split copy xh (x = hi + lo 2^-11), a row IS the
//         K = 64 operand row, and ONE [64 x 64] weight image per type (conv_t's: D[:, n] = hi.Whi,
//         D[:, 32+n] = hi.Wlo + lo.Whi) gives the pass in 4 MMAs (kind::f16, N = 64, K = 16).  Several same-type
//         in-edges of one destination (13 % of the (row, type) pairs on the synthetic graphs) are summed in fp32 and
//         re-split by the warp cooperatively (8 lanes per row) into a per-warp scratch row before the owner reads it.
//       k_conv_z<false> (stand-by, and TGNN_CONV=z32): fp32 rows, 3xTF32 (hi.Whi + lo.Whi + hi.Wlo, 12 MMAs of
//         N = 32, K = 8 per pass), full fp32 range.  It takes the layer when a range flag is raised (an activation,
//         a root weight or a multi-edge sum outside the fp16 range) -- launched behind the fast kernel, exits at once
//         otherwise.
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

#include "hsplit.cuh"
#include "layouts.cuh"
#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {
using namespace tc;

constexpr int Z_NG = 3;                          // gather groups of 4 warps
constexpr int Z_NB = 8;                          // weight images in flight (freed by the MMAs' commit)
constexpr int Z_NT = 16;                         // pass tables in flight (freed by the gather warps as soon as they are read)
// TMEM columns: two accumulator buffers, then the A operand stages.
//   fp16 : buffer = {agg main 32 | agg small 32 | root main 32 | root small 32}; A stage = 32 columns (64 halves); 8 stages
//   tf32 : buffer = {agg 32 | root 32};                                           A stage = 64 columns (hi | lo);   6 stages
template <bool HALF> struct ZCfg {
    static constexpr int D_COLS = HALF ? 128 : 64, ROOT_OFF = HALF ? 64 : 32, A_COLS = HALF ? 32 : 64, NSTA = HALF ? 8 : 6;
    static constexpr int A0 = 2 * D_COLS;
    static_assert(A0 + NSTA * A_COLS <= 512, "TMEM columns");
};
constexpr int Z_NSTA_MAX = 8;
constexpr float LO_INV = 1.0f / 2048.f;
constexpr int W_PRODW = 0, W_PRODT = 1, W_MMA = 2, W_PRODB = 3, W_G0 = 4, W_EPI0 = W_G0 + 4 * Z_NG;
constexpr int CZ_THREADS = (W_EPI0 + 4) * 32;
constexpr int SB_TILE = 4096;                    // bytes of one weight image (hi or lo): 32 rows x 128 B
// pass tables (one TAB_BYTES slot per pass):
constexpr int TAB_ZT = 0;                        // direct table: window row of each destination's (first) source, see k_zw_table
constexpr int TAB_OFF = TAB_ZT + 256;            // row-offset table of the pass (S_OFF_STRIDE uint16), read for multi-edge rows only
constexpr int TAB_LOC = TAB_OFF + 288;           // window offsets of the pass's edges (16-byte aligned superset), multi-edge rows only
constexpr int TAB_BYTES = 1024;
constexpr int ZT_NONE = 0xFFFF, ZT_MULTI = 0x8000;
static_assert(TAB_LOC + (ZW_MAX_PASS + 16) * 2 <= TAB_BYTES, "pass table slot too small");
static_assert(ZW_WMAX < ZT_MULTI, "window offsets need 15 bits");
constexpr int SCR_ROWS = 16;                     // re-split multi-edge rows per gather warp and pass (more -> the tf32 stand-by takes the layer)
constexpr int OFF_B = 0;                                        // [Z_NB][8 KB] weight images (1024-byte aligned)
constexpr int OFF_TAB = OFF_B + Z_NB * 2 * SB_TILE;             // [Z_NT][TAB_BYTES]
constexpr int OFF_WIN = OFF_TAB + Z_NT * TAB_BYTES;             // [ZW_WMAX][128 B]
constexpr int OFF_RED = OFF_WIN + ZW_WMAX * 128;                // [4][2][32] double: BatchNorm partial sums of the four epilogue warps
constexpr int OFF_SCR = OFF_RED + 4 * 2 * 32 * 8;               // [4 * Z_NG warps][SCR_ROWS][128 B]: re-split multi-edge rows (fp16 variant)
constexpr int CZ_SMEM = OFF_SCR + 4 * Z_NG * SCR_ROWS * 128 + 1024;
static_assert(CZ_SMEM + 1024 <= 232448, "shared memory (dynamic + the static barriers)");
// instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = 64
constexpr uint32_t IDESC_F16_N64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct ConvZArgs {
    const void* rows;            // [n_rows][128 B]: xh (fp16 variant) or xin (tf32 variant)
    const void* img;             // [K+1][8 KB] weight images: tabT (fp16 variant) or tabS (tf32 variant); entry K = root
    const int* flag_x; const int* flag_w; int* flag_z;   // range flags of the layer (flag_z: a multi-edge sum left the fp16 range)
    int force32;                 // TGNN_CONV=z32: the tf32 variant takes every layer
    int n_types;
    const int* pptr; const int* ptype; const int* pbase; const unsigned short* off;
    const int* zmeta; const int* zseg; const unsigned short* zloc; const unsigned short* ztab;
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
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ int4 sel4(bool p, const int4& a, const int4& b) {
    return make_int4(p ? a.x : b.x, p ? a.y : b.y, p ? a.z : b.z, p ? a.w : b.w);
}
__device__ __forceinline__ __half2 as_h2(int w) { return *reinterpret_cast<const __half2*>(&w); }
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
// a <- fl(a + b); returns the rounding error (a + b) - fl(a + b), exact in half precision (Knuth's TwoSum)
__device__ __forceinline__ __half2 two_sum(__half2& a, __half2 b) {
    const __half2 s = __hadd2(a, b), bb = __hsub2(s, a);
    const __half2 err = __hadd2(__hsub2(a, __hsub2(s, bb)), __hsub2(b, bb));
    a = s;
    return err;
}
__device__ __forceinline__ bool h2_nonfinite(uint32_t w) { return (w & 0x7C00u) == 0x7C00u || (w & 0x7C000000u) == 0x7C000000u; }

// Bounded mbarrier wait whose retry loop is three instructions (try_wait parks the warp until some mbarrier event or
// the time hint, whichever comes first -- with many barriers in flight it returns every few dozen cycles): the clock
// is only looked at every 64K retries.  `backoff` (ns) for roles off the critical path.
__device__ __forceinline__ bool zwait(uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t it = 1;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (ok) return true;
        if ((it & 0xFFFFu) == 0) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000ll) return false;
        }
    }
}
// Roles with slack (ring producers, epilogue): plain polling with a real sleep between probes.  try_wait's parked warp
// is woken by EVERY mbarrier event of the CTA (measured: ~28 wake-ups per pass and warp), which is fine for the roles on
// the critical path and wasteful for the others.
__device__ __forceinline__ bool zwait_lazy(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
    long long t0 = 0;
    for (uint32_t it = 1;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        __nanosleep(sleep_ns);
        if ((it & 0xFFFu) == 0) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 4000000000ll) return false;
        }
    }
}

template <bool HALF>
__global__ void __launch_bounds__(CZ_THREADS, 1)
k_conv_z(ConvZArgs A) {
    using Z = ZCfg<HALF>;
    {
        const bool flagged = A.force32 || (A.flag_x && *A.flag_x) || (A.flag_w && *A.flag_w) || (A.flag_z && *A.flag_z);
        if (HALF ? flagged : !flagged) return;          // (a flag_z raised by a sibling CTA mid-kernel only ends this one early:
    }                                                   //  the stand-by recomputes the whole layer)
    // role timing only when asked for (TGNN_ROLE_DBG=1): clock reads cost issue slots in the gather warps.
    // All waits use mbarrier.try_wait with a suspend hint (the hardware parks the warp): a nanosleep polling loop wakes
    // up every ~80 cycles and its instructions compete with the gather warps, which are issue-bound (ncu: the four
    // epilogue warps polling executed 37 % of all instructions of the kernel).
    long long w0 = 0, w1 = 0, w2 = 0;                  // cycles spent in this role's barrier waits
    const bool timed = A.dbg != nullptr;
#define ZT(acc, expr) (timed ? TGNN_TIMED(acc, expr) : (expr))
    const long long t_start = timed ? clock64() : 0;
    extern __shared__ uint8_t smem_raw[];
    // win_full, win_empty, tab_full[NT], tab_empty[NT], b_full[NB], b_empty[NB], a_full[NSTA], a_empty[NSTA], acc_full[2], acc_empty[2]
    __shared__ __align__(8) uint64_t bars[2 + 2 * Z_NT + 2 * Z_NB + 2 * Z_NSTA_MAX + 4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int timeout_flag;
    __shared__ int tab_shift[Z_NT];                    // first edge of the pass inside the slot's (aligned) offset copy
    __shared__ int win_self;                           // window row of the tile's first own row
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_wf = smem_u32(&bars[0]), bar_we = smem_u32(&bars[1]);
    const uint32_t bar_tf = smem_u32(&bars[2]), bar_te = bar_tf + 8 * Z_NT;
    const uint32_t bar_bf = bar_te + 8 * Z_NT, bar_be = bar_bf + 8 * Z_NB;
    const uint32_t bar_af = bar_be + 8 * Z_NB, bar_ae = bar_af + 8 * Z_NSTA_MAX;
    const uint32_t bar_cf = bar_ae + 8 * Z_NSTA_MAX, bar_ce = bar_cf + 16;
    constexpr int Z_NSTA = Z::NSTA;

    if (tid == 0) {
        mbar_init(bar_wf, 1); mbar_init(bar_we, 4 * Z_NG);
        for (int i = 0; i < Z_NT; ++i) { mbar_init(bar_tf + 8 * i, 1); mbar_init(bar_te + 8 * i, 4); }          // freed by the pass's 4 gather warps
        for (int i = 0; i < Z_NB; ++i) { mbar_init(bar_bf + 8 * i, 1); mbar_init(bar_be + 8 * i, 1); }          // freed by the MMA commit
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
            if (!ZT(w0, zwait(bar_we, (uint32_t)((it & 1) ^ 1)))) { timeout_flag = 1; break; }
            if (lane == 0) {
                win_self = m.z;
                mbar_arrive_expect_tx(bar_wf, (uint32_t)m.y * 128u);      // (release: win_self is visible to the waiters)
            }
            __syncwarp();
            if (lane < m.x)
                bulk_g2s(sbase + OFF_WIN + (uint32_t)sg.y * 128u, reinterpret_cast<const char*>(A.rows) + (size_t)sg.x * 128,
                         (uint32_t)(next - sg.y) * 128u, bar_wf);
        }
    } else if (warp == W_PRODT || warp == W_PRODB) {
        // ===================== pass producers: warp 1 the tables (deep ring, the gather warps run ahead of the MMAs),
        // ===================== warp 3 the weight images =====================
        const bool tabs = warp == W_PRODT;
        int s = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x) {
            const int p0 = __ldg(A.pptr + tile), np = __ldg(A.pptr + tile + 1) - p0;
            for (int qb = 0; qb <= np && ok; qb += 32) {
                // lane l looks up pass qb + l; lane 0 issues the copies in pass order
                const int q = qb + lane;
                int type = A.n_types, eb = 0, ee = 0;
                if (q < np) {
                    if (tabs) { eb = __ldg(A.pbase + p0 + q); ee = __ldg(A.pbase + p0 + q + 1); }
                    else type = __ldg(A.ptype + p0 + q);
                }
                const int nq = min(32, np + 1 - qb);
                for (int j = 0; j < nq; ++j, ++s) {
                    const int ty = __shfl_sync(0xffffffffu, type, j), b0 = __shfl_sync(0xffffffffu, eb, j), b1 = __shfl_sync(0xffffffffu, ee, j);
                    const bool root = qb + j == np;
                    if (tabs) {
                        const int slot = s % Z_NT;
                        if (!ZT(w0, zwait_lazy(bar_te + 8 * slot, (uint32_t)(((s / Z_NT) & 1) ^ 1), 600))) { ok = false; break; }
                        if (lane == 0) {
                            const uint32_t dst = sbase + OFF_TAB + (uint32_t)slot * TAB_BYTES, bar = bar_tf + 8 * slot;
                            if (root) mbar_arrive(bar);                                    // the root pass has no tables: phase completes at once
                            else {
                                const int a0 = b0 & ~7, a1 = (b1 + 7) & ~7;
                                tab_shift[slot] = b0 - a0;
                                const uint32_t loc_bytes = (uint32_t)(a1 - a0) * 2u;
                                mbar_arrive_expect_tx(bar, 256u + (uint32_t)S_OFF_STRIDE * 2u + loc_bytes);
                                bulk_g2s(dst + TAB_ZT, A.ztab + (size_t)(p0 + qb + j) * 128, 256u, bar);
                                bulk_g2s(dst + TAB_OFF, A.off + (size_t)(p0 + qb + j) * S_OFF_STRIDE, (uint32_t)S_OFF_STRIDE * 2u, bar);
                                if (loc_bytes) bulk_g2s(dst + TAB_LOC, A.zloc + a0, loc_bytes, bar);
                            }
                        }
                    } else {
                        const int slot = s % Z_NB;
                        if (!ZT(w0, zwait_lazy(bar_be + 8 * slot, (uint32_t)(((s / Z_NB) & 1) ^ 1), 400))) { ok = false; break; }
                        if (lane == 0) {
                            const uint32_t bar = bar_bf + 8 * slot;
                            mbar_arrive_expect_tx(bar, 2u * SB_TILE);
                            bulk_g2s(sbase + OFF_B + (uint32_t)slot * (2 * SB_TILE), reinterpret_cast<const char*>(A.img) + (size_t)ty * (2 * SB_TILE),
                                     2u * SB_TILE, bar);
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
            const uint64_t dB = umma_desc_sw128(sbase + OFF_B);
            int s = 0, it = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x, ++it) {
                const int np = __ldg(A.pptr + tile + 1) - __ldg(A.pptr + tile);
                const int ab = it & 1;
                if (!ZT(w2, zwait(bar_ce + 8 * ab, (uint32_t)(((it >> 1) & 1) ^ 1)))) { ok = false; break; }
                for (int q = 0; q <= np; ++q, ++s) {
                    const bool root = q == np;
                    const int slot = s % Z_NB, stg = s % Z_NSTA;
                    if (!ZT(w0, zwait(bar_bf + 8 * slot, (uint32_t)((s / Z_NB) & 1)))) { ok = false; break; }
                    if (!ZT(w1, zwait(bar_af + 8 * stg, (uint32_t)((s / Z_NSTA) & 1)))) { ok = false; break; }
                    fence_proxy_async();
                    tc_fence_after();
                    const uint64_t dbh = dB + (uint64_t)(slot * (2 * SB_TILE / 16)), dbl = dbh + SB_TILE / 16;
                    const uint32_t a_hi = tmem_base + (uint32_t)(Z::A0 + stg * Z::A_COLS), a_lo = a_hi + 32u;
                    const uint32_t tmem_d = tmem_base + (uint32_t)(ab * Z::D_COLS) + (root ? (uint32_t)Z::ROOT_OFF : 0u);
                    const uint32_t first = (root || q == 0) ? 0u : 1u;
                    if (HALF) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)     // 16 halves = 8 TMEM columns of A = 32 bytes = 2 descriptor units of B inside the swizzle atom
                            umma_f16_ts(tmem_d, a_hi + 8 * ks, dbh + 2 * ks, IDESC_F16_N64, ks == 0 ? first : 1u);
                    } else {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {   // 8 tf32 = 8 TMEM columns of A = 32 bytes = 2 descriptor units of B
                            umma_tf32_ts(tmem_d, a_lo + 8 * ks, dbh + 2 * ks, IDESC, ks == 0 ? first : 1u);
                            umma_tf32_ts(tmem_d, a_hi + 8 * ks, dbl + 2 * ks, IDESC, 1u);
                            umma_tf32_ts(tmem_d, a_hi + 8 * ks, dbh + 2 * ks, IDESC, 1u);
                        }
                    }
                    umma_commit(bar_be + 8 * slot);        // weight image free
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
        const uint32_t scr = sbase + OFF_SCR + (uint32_t)(warp - W_G0) * (SCR_ROWS * 128);
        uint32_t co[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) co[k] = (uint32_t)(((k + rho) & 7) << 4);
        int s = 0, it = 0;
        bool ok = true, ovf = false;
        for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x, ++it) {
            const int np = __ldg(A.pptr + tile + 1) - __ldg(A.pptr + tile);
            const bool live = tile * S_BM + r < A.n_own;
            if (!ZT(w2, zwait(bar_wf, (uint32_t)(it & 1)))) { ok = false; break; }
            const int self_loc = win_self;
            const int s_tile = s;                                  // pass counter of the tile's first pass
            s += np + 1;
            for (int q = (grp + Z_NG - s_tile % Z_NG) % Z_NG; q <= np; q += Z_NG) {      // this group's passes: (s_tile + q) % NG == grp
                const int s = s_tile + q;
                const bool root = q == np;
                const int slot = s % Z_NT, stg = s % Z_NSTA;
                if (!ZT(w0, zwait(bar_tf + 8 * slot, (uint32_t)((s / Z_NT) & 1)))) { ok = false; break; }
                const uint32_t sl = sbase + OFF_TAB + (uint32_t)slot * TAB_BYTES;
                // direct table: window row of the destination's source | ZT_MULTI when it has several of this type (then the
                // count and the list come from the offset tables), ZT_NONE when it has none
                const uint32_t zt = root ? (live ? (uint32_t)(self_loc + r) : (uint32_t)ZT_NONE) : lds_u16(sl + TAB_ZT + 2 * r);
                int cnt = zt == (uint32_t)ZT_NONE ? 0 : 1;
                uint32_t lp = 0;
                if (cnt && (zt & ZT_MULTI)) {
                    const int e0 = (int)lds_u16(sl + TAB_OFF + 2 * r), e1 = (int)lds_u16(sl + TAB_OFF + 2 * r + 2);
                    cnt = e1 - e0;
                    lp = sl + TAB_LOC + 2u * (uint32_t)(tab_shift[slot] + e0);
                }
                int4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = make_int4(0, 0, 0, 0);
                uint32_t rowa = win + (zt & (uint32_t)(ZT_MULTI - 1)) * 128u;
                if (HALF) {
                    // several same-type in-edges of one destination: the split parts cannot be added as they are, so the warp
                    // sums those rows in fp32 (x = hi + lo 2^-11 is exact), splits the sum again and parks it in its scratch
                    // rows -- four rows per round, 8 lanes (one 16-byte piece = 4 channels each) per row
                    unsigned multi = __ballot_sync(0xffffffffu, cnt > 1);
                    const unsigned multi0 = multi;                 // scratch row of a multi lane = its rank among them
                    if (__popc(multi0) > SCR_ROWS) { ovf = true; multi = 0; }   // (does not happen on tile graphs: the stand-by redoes the layer)
                    const int ga = lane >> 3, qc = lane & 7;
                    while (multi) {
                        unsigned m = multi;
                        int owner = -1;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (m) { const int b = __ffs(m) - 1; if (j == ga) owner = b; m &= m - 1; }
                        multi = m;
                        const int sl_lane = owner >= 0 ? owner : 0;
                        const int ocnt = __shfl_sync(0xffffffffu, cnt, sl_lane);
                        const uint32_t olp = __shfl_sync(0xffffffffu, lp, sl_lane);
                        if (owner >= 0) {
                            // error-free sum of the split parts in half precision (no unpacking, no re-split):
                            //   hi' = fl(hi_a + hi_b),  err = (hi_a + hi_b) - hi' exactly (TwoSum),  lo' = fl(lo_a + lo_b + 2^11 err)
                            // keeps x_a + x_b to ~2^-22 relative, like the split itself
                            const int4 w = lds128i(win + lds_u16(olp) * 128u + (uint32_t)qc * 16u);
                            __half2 h0 = as_h2(w.x), h1 = as_h2(w.y), l0 = as_h2(w.z), l1 = as_h2(w.w);
                            const __half2 k2048 = __float2half2_rn(2048.f);
                            for (int t = 1; t < ocnt; ++t) {
                                const int4 y = lds128i(win + lds_u16(olp + 2u * (uint32_t)t) * 128u + (uint32_t)qc * 16u);
                                l0 = __hfma2(two_sum(h0, as_h2(y.x)), k2048, __hadd2(l0, as_h2(y.z)));
                                l1 = __hfma2(two_sum(h1, as_h2(y.y)), k2048, __hadd2(l1, as_h2(y.w)));
                            }
                            const uint4 sp = make_uint4(h2_bits(h0), h2_bits(h1), h2_bits(l0), h2_bits(l1));
                            ovf |= h2_nonfinite(sp.x) | h2_nonfinite(sp.y) | h2_nonfinite(sp.z) | h2_nonfinite(sp.w);
                            sts128(scr + (uint32_t)__popc(multi0 & ((1u << owner) - 1u)) * 128u + (uint32_t)qc * 16u, sp);
                        }
                    }
                    __syncwarp();
                    if (cnt > 1) rowa = scr + (uint32_t)(__popc(multi0 & ((1u << lane) - 1u)) & (SCR_ROWS - 1)) * 128u;
                }
                if (cnt > 0) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = lds128i(rowa + co[k]);
                }
                if (!HALF) {
                    for (int t = 1; t < cnt; ++t) {       // multi-edges of one (row, type): fp32 sum, like the reference's scatter
                        const uint32_t ra = win + lds_u16(lp + 2u * (uint32_t)t) * 128u;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float4 w = lds128f(ra + co[k]);
                            v[k].x = __float_as_int(__int_as_float(v[k].x) + w.x); v[k].y = __float_as_int(__int_as_float(v[k].y) + w.y);
                            v[k].z = __float_as_int(__int_as_float(v[k].z) + w.z); v[k].w = __float_as_int(__int_as_float(v[k].w) + w.w);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_te + 8 * slot);     // the pass's tables are read: the slot goes back to the producer
                // v[k] holds chunk (k + rho) & 7 of the row: rotate back by rho in three select stages
                int4 a[8], b[8], u[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) a[c] = sel4(rho & 1, v[(c + 7) & 7], v[c]);
#pragma unroll
                for (int c = 0; c < 8; ++c) b[c] = sel4(rho & 2, a[(c + 6) & 7], a[c]);
#pragma unroll
                for (int c = 0; c < 8; ++c) u[c] = sel4(rho & 4, b[(c + 4) & 7], b[c]);
                // the A stage was last read by the MMAs of pass s - NSTA
                if (!ZT(w1, zwait(bar_ae + 8 * stg, (uint32_t)(((s / Z_NSTA) & 1) ^ 1)))) { ok = false; break; }
                tc_fence_after();
                const uint32_t ta = tmem_base + t_lane + (uint32_t)(Z::A0 + stg * Z::A_COLS);
                uint32_t hi[32];
                if (HALF) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {              // the split row in storage order IS the K = 64 operand row
                        hi[4 * c + 0] = (uint32_t)u[c].x; hi[4 * c + 1] = (uint32_t)u[c].y; hi[4 * c + 2] = (uint32_t)u[c].z; hi[4 * c + 3] = (uint32_t)u[c].w;
                    }
                    tmem_st32(ta, hi);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        hi[4 * c + 0] = (uint32_t)u[c].x & 0xFFFFE000u; hi[4 * c + 1] = (uint32_t)u[c].y & 0xFFFFE000u;
                        hi[4 * c + 2] = (uint32_t)u[c].z & 0xFFFFE000u; hi[4 * c + 3] = (uint32_t)u[c].w & 0xFFFFE000u;
                    }
                    tmem_st32(ta, hi);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {              // lo = x - hi is exact; the tensor core truncates it to TF32 itself
                        hi[4 * c + 0] = __float_as_uint(__int_as_float(u[c].x) - __uint_as_float(hi[4 * c + 0]));
                        hi[4 * c + 1] = __float_as_uint(__int_as_float(u[c].y) - __uint_as_float(hi[4 * c + 1]));
                        hi[4 * c + 2] = __float_as_uint(__int_as_float(u[c].z) - __uint_as_float(hi[4 * c + 2]));
                        hi[4 * c + 3] = __float_as_uint(__int_as_float(u[c].w) - __uint_as_float(hi[4 * c + 3]));
                    }
                    tmem_st32(ta + 32u, hi);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_af + 8 * stg);
            }
            if (!ok) break;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_we);                    // this warp is done with the tile's window
        }
        if (HALF && __any_sync(0xffffffffu, ovf) && lane == 0 && A.flag_z) *A.flag_z = 1;   // the tf32 stand-by redoes the layer
        if (!ok) timeout_flag = 1;
    } else if (warp >= W_EPI0) {
        // ===================== epilogue warps: TMEM lane quarter q4 = warp % 4 =====================
        const int q4 = warp & 3, etid = (warp - W_EPI0) * 32 + lane;
        const uint32_t red = sbase + OFF_RED;
        int it = 0;
        for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const int np = __ldg(A.pptr + tile + 1) - __ldg(A.pptr + tile);
            if (!ZT(w0, zwait_lazy(bar_cf + 8 * ab, (uint32_t)((it >> 1) & 1), 1000))) { timeout_flag = 1; break; }
            tc_fence_after();
            const int row = tile * S_BM + 32 * q4 + lane;
            const bool live = row < A.n_own;
            const float idg = live ? __ldg(A.inv_deg + row) : 0.f;
            const bool kept = live && row_kept(A.mask, row);
            const uint32_t tbase = tmem_base + ((uint32_t)(32 * q4) << 16) + (uint32_t)(ab * Z::D_COLS);
            float o[32];
            {
                uint32_t vt[32], vr[32];
                if (np > 0) {
                    tmem_ld32(tbase, vt);
                    if (HALF) {
                        tmem_ld32(tbase + 32u, vr);
#pragma unroll
                        for (int j = 0; j < 32; ++j) vt[j] = __float_as_uint(fmaf(__uint_as_float(vr[j]), LO_INV, __uint_as_float(vt[j])));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) vt[j] = 0u;
                }
                tmem_ld32(tbase + (uint32_t)Z::ROOT_OFF, vr);
#pragma unroll
                for (int j = 0; j < 32; ++j) o[j] = fmaf(__uint_as_float(vt[j]), idg, __uint_as_float(vr[j]));
                if (HALF) {
                    tmem_ld32(tbase + (uint32_t)Z::ROOT_OFF + 32u, vr);
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = fmaf(__uint_as_float(vr[j]), LO_INV, o[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ce + 8 * ab);          // accumulator buffer free for the tile after next
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = kept ? leaky(o[j] + __ldg(A.bias + j)) : 0.f;
            if (live) {
                float4* dst = reinterpret_cast<float4*>(A.out + (size_t)row * F);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            if (A.part) {
                // column sums over the warp's 32 rows in fp64 by recursive halving (fixed order -> bit-reproducible): 16 channels
                // at a time, first across lane bit 4, then each step halves the channels a lane carries; lane l ends with
                // channel 16 h + (l & 15).  (Rows past n_own and masked rows hold 0.)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                    for (int st = 0; st < 2; ++st) {
                        double a[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) { const double x = (double)o[16 * hf + i]; a[i] = st ? x * x : x; }
#pragma unroll
                        for (int i = 0; i < 16; ++i) a[i] += __shfl_xor_sync(0xffffffffu, a[i], 16);
#pragma unroll
                        for (int half = 8; half >= 1; half >>= 1) {
                            const bool up = (lane & half) != 0;
#pragma unroll
                            for (int i = 0; i < half; ++i) {
                                const double keep = up ? a[i + half] : a[i], send = up ? a[i] : a[i + half];
                                a[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
                            }
                        }
                        if (lane < 16) sts_f64(red + 8 * ((q4 * 2 + st) * 32 + 16 * hf + lane), a[0]);
                    }
                }
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

// direct table: per (pass, destination row) the window row of the row's first source of that type, | ZT_MULTI when there
// are several, ZT_NONE when there is none -- the common case costs the gather thread ONE shared-memory load
__global__ void __launch_bounds__(128)
k_zw_table(const int* __restrict__ pbase, const unsigned short* __restrict__ off, const unsigned short* __restrict__ loc,
           unsigned short* __restrict__ ztab) {
    const int p = blockIdx.x, r = threadIdx.x;
    const int e0 = off[(size_t)p * S_OFF_STRIDE + r], e1 = off[(size_t)p * S_OFF_STRIDE + r + 1];
    ztab[(size_t)p * 128 + r] = e1 > e0 ? (unsigned short)(loc[pbase[p] + e0] | (e1 - e0 > 1 ? ZT_MULTI : 0)) : (unsigned short)ZT_NONE;
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
    if (n_bad == 0 && g.s_passes > 0) {
        g.z_tab.reserve((size_t)g.s_passes * 128 * sizeof(unsigned short));
        k_zw_table<<<g.s_passes, 128, 0, st>>>(g.s_pbase.as<int>(), g.s_off.as<unsigned short>(), g.z_loc.as<unsigned short>(),
                                               g.z_tab.as<unsigned short>());
        TGNN_CUDA(cudaGetLastError());
    }
    return n_bad;
}

int conv_z_blocks(int s_tiles, int sm_count) { return s_tiles < sm_count ? (s_tiles < 1 ? 1 : s_tiles) : sm_count; }

// Two launches: the fp16 kernel, and the tf32 stand-by that exits at once unless a range flag is raised (or force32).
void launch_conv_z(const ConvArgs& c, const Graph& g, const float* tabS, const uint32_t* tabT, int* flag_z, bool force32, int* error_flag,
                   int sm_count, cudaStream_t st, long long* dbg) {
    static PerDeviceOnce once;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_z<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CZ_SMEM));
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_z<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CZ_SMEM));
    });
    TGNN_CHECK(g.has_z, "internal: k_conv_z without its windows");
    ConvZArgs a{};
    a.flag_x = c.flag_x; a.flag_w = c.flag_w; a.flag_z = flag_z; a.force32 = force32 ? 1 : 0;
    a.n_types = g.n_types;
    a.pptr = g.s_pptr.as<int>(); a.ptype = g.s_ptype.as<int>(); a.pbase = g.s_pbase.as<int>();
    a.off = g.s_off.as<unsigned short>();
    a.zmeta = g.zw_meta.as<int>(); a.zseg = g.zw_seg.as<int>(); a.zloc = g.z_loc.as<unsigned short>(); a.ztab = g.z_tab.as<unsigned short>();
    a.inv_deg = c.inv_deg; a.bias = c.bias; a.out = c.out; a.part = c.part; a.error_flag = error_flag; a.mask = c.mask;
    a.n_own = c.n_own; a.n_tiles = g.s_tiles; a.dbg = dbg;
    const int blocks = conv_z_blocks(g.s_tiles, sm_count);
    if (!force32) {
        a.rows = c.xh; a.img = tabT;
        k_conv_z<true><<<blocks, CZ_THREADS, CZ_SMEM, st>>>(a);
        TGNN_CUDA(cudaGetLastError());
    }
    a.rows = c.xin; a.img = tabS;
    k_conv_z<false><<<blocks, CZ_THREADS, CZ_SMEM, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
