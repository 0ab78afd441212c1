// Adjacency branch on the 5th-generation tensor cores, EDGE-BLOCK formulation ("T" kernel): typed NNConv(mean) + root +
// bias + LeakyReLU and the BatchNorm partial sums (graph_networks/layers/edge_conv.py:24-27 of the reference; PyG
// NNConv semantics) -- the same arithmetic as k_conv_h (fp16 two-term split operands, 2^-22 relative), but nothing of
// the per-edge product is executed by the SM's warps any more:
//   * the A operand of one tcgen05.mma (M = 128) is a BLOCK of 128 same-type edges: their split rows (xh: 64 fp16 =
//     128 B = exactly one SWIZZLE_128B row, written by the producers of b1) are gathered straight into the swizzled
//     operand tile by six producer warps: 16-byte loads into REGISTERS (8 lanes per 128-byte row = whole lines per
//     request), three blocks of loads in flight per thread (~55 KB per SM), then swizzled 128-bit stores into the stage --
//     the registers are the latency buffer, the loads run ahead of the shared-memory stages.  Measured alternatives on
//     B200, all bound by the gather: cp.async (LDGSTS) keeps only ~8 copies per WARP outstanding (four warps per block:
//     ~1000 cycles per block; one warp per block: ~9000); TMA tile::gather4 (UTMALDG.2D.GATHER4: correct, but ~75 cycles
//     of TMA service per 512-byte instruction = 2400 cycles per block).
//   * the B operand is the type's pre-swizzled [64 x 64] fp16 image (tables.cu), one TMA bulk copy per block:
//       columns  0..31 : hi . Whi                       ("main")
//       columns 32..63 : hi . Wlo + lo . Whi            ("small", scaled by 2^-11 in the epilogue)
//     so ONE chain of four tcgen05.mma.kind::f16 (M = 128, N = 64, K = 16) per block replaces the 24 mma.sync of a
//     16-edge chunk x 8, the A tile is read from shared memory once, and the accumulator lives in TMEM;
//   * the epilogue warps only read their edge's 64 accumulator columns (tcgen05.ld), combine, and add the message into
//     the destination row of the super-tile's accumulator in shared memory.  Thread = edge, so the float4 pieces of a
//     row are visited in a lane-rotated order (conflict-free for ANY destinations); the data is rotated to match with
//     three select stages.  Quarter q of a block only holds destinations with dst % 4 == q and no destination twice
//     (graph_build.cu), so the four warps never touch the same row: no atomics, deterministic sums.
//   * x_i . root runs as the super-tile's last blocks (type K, src = dst = own rows); edge messages are scaled by 1 / deg(dst)
//     as they are added (mean aggregation), the root message is added as it is.
//   * measured (role cycle counters): one epilogue warp per TMEM lane quarter is the bottleneck (~900 busy cycles per block,
//     a single warp per scheduler cannot hide its own tcgen05.ld -> FMA -> select -> FADD -> STS chain).  So there are TWO
//     epilogue groups of four warps; group e takes the blocks with (block counter & 1) == e and accumulates into its OWN
//     plane of the super-tile accumulator (two groups adding into one plane would race on rows of the same class); the
//     planes are added when the tile is finished.  The accumulator rows are fetched BEFORE the wait for the MMAs.
// Persistent CTAs (one per SM), 13 warps: 0-7 epilogue (warp & 3 = TMEM lane quarter, warp >> 2 = group), 8-11 gather
// producers, 12 MMA issuer.  Super-tiles of 512 rows: 2 x 64 KB of accumulator planes + 4 stages of 24 KB.
// Range guard: fp16 operands overflow at 65504; when a range flag is raised (an activation or a root weight above
// 60000) the kernel exits at once and k_conv_t_wide redoes the layer on the same blocks with plain fp32 FMAs.
#include <algorithm>
#include <cstdlib>

#include "hsplit.cuh"
#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {
using namespace tc;

constexpr int TBS = 128;                        // slots (edges) per block = UMMA M
constexpr int NS = 4;                           // shared-memory stages (A tile + weight image)
constexpr int NT = 4;                           // TMEM accumulator buffers (64 columns each)
constexpr int A_BYTES = TBS * 128, B_BYTES = 64 * 128, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_GROUPS = 2, EPI_WARPS = 4 * EPI_GROUPS, PROD_WARPS = 6, PROD_THREADS = PROD_WARPS * 32;
constexpr int W_PROD0 = EPI_WARPS, W_MMA = EPI_WARPS + PROD_WARPS;
constexpr int CT_THREADS = (W_MMA + 1) * 32;                     // 15 warps
constexpr int CPB = (TBS * 8 + PROD_THREADS - 1) / PROD_THREADS;   // 16-byte pieces per producer thread per block (6)
constexpr int PDEPTH = 3;                                       // blocks of row loads in flight per producer thread
constexpr float LO_INV = 1.0f / 2048.f;

// instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = 64
constexpr uint32_t IDESC_F16_N64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on the mbarrier once all of this thread's earlier cp.async have landed (does not change the pending count)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
                   "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),
                   "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),
                   "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct ConvTArgs {
    const uint4* xh;             // [n_rows][8] fp16-split rows (hsplit.cuh)
    const float* xin;            // [n_rows][32] fp32 rows (k_conv_t_wide)
    const uint32_t* tabT;        // [K+1][2048] words: pre-swizzled [64 x 64] fp16 weight images (entry K = root)
    const float* tab32;          // [K+1][1024] fp32 weights [k_in][k_out] (k_conv_t_wide)
    int n_types;
    const int* bptr; const int* btype; const int* tsrc; const unsigned short* tdst;
    const float* inv_deg; const float* bias;
    float* out; double* part;
    const int* flag_x; const int* flag_w; int* err;
    const uint8_t* mask;
    long long* dbg;              // optional: per-warp {cycles, wait 0, wait 1, wait 2} of CTA 0 (TGNN_ROLE_DBG=1)
    int n_own, n_tiles, rt;      // rt: destination rows per super-tile (256 | 512 | 1024)
};

// out = LeakyReLU(plane 0 + plane 1 + bias), statistics, and the tile's accumulator planes are zeroed for the next tile.
// Called by the eight epilogue warps (warp w takes rows w, w+8, ...; lane = channel) between two named barriers.
__device__ __forceinline__ void finish_tile(const ConvTArgs& A, uint32_t acc_base, int tile, int w, int lane, float bias_c,
                                            double& s1, double& s2) {
    const int node0 = tile * A.rt;
    const int rows = min(A.rt, A.n_own - node0);
    const uint32_t plane = (uint32_t)A.rt * 128u;
    for (int r = w; r < rows; r += EPI_WARPS) {
        const uint32_t a = acc_base + (uint32_t)r * 128u + (uint32_t)lane * 4u;
        float v = leaky((lds_f32(a) + lds_f32(a + plane)) + bias_c);
        if (!row_kept(A.mask, node0 + r)) v = 0.f;
        sts_f32(a, 0.f);
        sts_f32(a + plane, 0.f);
        A.out[(size_t)(node0 + r) * F + lane] = v;
        s1 += (double)v;
        s2 += (double)v * (double)v;
    }
}

__global__ void __launch_bounds__(CT_THREADS, 1)
k_conv_t(ConvTArgs A) {
    if ((A.flag_x && *A.flag_x) || (A.flag_w && *A.flag_w)) return;          // out of the fp16 range: k_conv_t_wide takes the layer
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * NS + 2 * NT];                  // full[NS], empty[NS], accf[NT], acce[NT]
    __shared__ uint32_t tmem_base_smem;
    __shared__ int timeout_flag;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t acc_base = sbase;                                          // [2 planes][rt][32] fp32, row stride 128 B
    const uint32_t stage_base = sbase + (uint32_t)EPI_GROUPS * (uint32_t)A.rt * 128u;   // [NS][A tile | weight image]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[NS]);
    const uint32_t bar_accf = smem_u32(&bars[2 * NS]), bar_acce = smem_u32(&bars[2 * NS + NT]);
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(bar_full + 8 * i, PROD_THREADS + 1); mbar_init(bar_empty + 8 * i, 1); }   // the producers + the expect_tx
        for (int i = 0; i < NT; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_acce + 8 * i, 4); }     // the 4 warps of the owning group
        timeout_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(NT * 64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (uint32_t i = tid; i < (uint32_t)EPI_GROUPS * (uint32_t)A.rt * 8u; i += CT_THREADS) sts128f(acc_base + i * 16u, make_float4(0.f, 0.f, 0.f, 0.f));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    long long w0 = 0, w1 = 0, w2 = 0;
    const long long t_start = clock64();

    if (warp >= W_PROD0 && warp < W_MMA) {
        // ===================== producers: 16-byte piece id = k * 192 + p of the block's 128 x 8 pieces -> row id >> 3, piece id & 7 =====================
        // Per tile: slot j of the register ring holds the pieces of block j, j + 3, ...; a block is retired (stored into its
        // stage, fence, arrive) and the slot refilled with the loads of the block three ahead, whose source rows were fetched
        // three blocks before that.
        const int p = (warp - W_PROD0) * 32 + lane;
        int g = 0;
        bool ok = true;
        for (int tile = blockIdx.x; ok && tile < A.n_tiles; tile += gridDim.x) {
            const int b0 = __ldg(A.bptr + tile), n = __ldg(A.bptr + tile + 1) - b0;
            const int* tsrc = A.tsrc + (size_t)b0 * TBS;
            uint4 buf[PDEPTH][CPB];
            int idx[PDEPTH][CPB], type[PDEPTH];
            auto load_idx = [&](int (&ix)[CPB], int& ty, int blk) {
#pragma unroll
                for (int k = 0; k < CPB; ++k) {
                    const int id = k * PROD_THREADS + p;
                    ix[k] = id < TBS * 8 ? __ldg(tsrc + (size_t)blk * TBS + (id >> 3)) : -1;
                }
                ty = __ldg(A.btype + b0 + blk);
            };
            auto load_rows = [&](uint4 (&bf)[CPB], const int (&ix)[CPB]) {
#pragma unroll
                for (int k = 0; k < CPB; ++k) {
                    bf[k] = make_uint4(0u, 0u, 0u, 0u);                       // empty slot: a zero row (its result is never used)
                    if (ix[k] >= 0) bf[k] = __ldg(A.xh + (size_t)ix[k] * 8 + ((k * PROD_THREADS + p) & 7));
                }
            };
#pragma unroll
            for (int j = 0; j < PDEPTH; ++j) if (j < n) load_idx(idx[j], type[j], j);
            int cur_type[PDEPTH];                          // type of the block whose pieces sit in slot j
#pragma unroll
            for (int j = 0; j < PDEPTH; ++j) {
                cur_type[j] = 0;
                if (j < n) { load_rows(buf[j], idx[j]); cur_type[j] = type[j]; if (j + PDEPTH < n) load_idx(idx[j], type[j], j + PDEPTH); }
            }
            for (int base = 0; ok && base < n; base += PDEPTH) {
#pragma unroll
                for (int j = 0; j < PDEPTH; ++j) {
                    const int i = base + j;
                    if (i >= n) break;
                    const int s = g % NS;
                    if (!TGNN_TIMED(w0, mbar_wait_relaxed(bar_empty + 8 * s, (uint32_t)(((g / NS) & 1) ^ 1)))) { timeout_flag = 1; ok = false; break; }
                    const uint32_t a_tile = stage_base + (uint32_t)s * STAGE_BYTES, bar = bar_full + 8 * s;
                    if (p == 0) {
                        mbar_arrive_expect_tx(bar, B_BYTES);
                        bulk_g2s(a_tile + A_BYTES, A.tabT + (size_t)cur_type[j] * (B_BYTES / 4), B_BYTES, bar);
                    }
#pragma unroll
                    for (int k = 0; k < CPB; ++k) {
                        const int id = k * PROD_THREADS + p;
                        if (id < TBS * 8) sts128(a_tile + sw128_off(id >> 3, id & 7), buf[j][k]);
                    }
                    fence_proxy_async();                                      // generic-proxy stores -> visible to the tensor core's reads
                    mbar_arrive(bar);
                    ++g;
                    if (i + PDEPTH < n) {
                        load_rows(buf[j], idx[j]);                            // idx[j] holds the rows of block i + PDEPTH
                        cur_type[j] = type[j];
                        if (i + 2 * PDEPTH < n) load_idx(idx[j], type[j], i + 2 * PDEPTH);
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ===================== MMA issuer: four tcgen05.mma per block into TMEM buffer g % NT =====================
        int g = 0;
        bool ok = true;
        for (int tile = blockIdx.x; ok && tile < A.n_tiles; tile += gridDim.x) {
            const int b0 = __ldg(A.bptr + tile), b1 = __ldg(A.bptr + tile + 1);
            for (int blk = b0; blk < b1; ++blk, ++g) {
                const int s = g % NS, tb = g % NT;
                if (!TGNN_TIMED(w0, mbar_wait(bar_acce + 8 * tb, (uint32_t)(((g / NT) & 1) ^ 1)))) { timeout_flag = 1; ok = false; break; }
                if (!TGNN_TIMED(w1, mbar_wait(bar_full + 8 * s, (uint32_t)((g / NS) & 1)))) { timeout_flag = 1; ok = false; break; }
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_tile = stage_base + (uint32_t)s * STAGE_BYTES;
                    const uint64_t da = umma_desc_sw128(a_tile), db = umma_desc_sw128(a_tile + A_BYTES);
                    const uint32_t d = tmem_base + (uint32_t)(tb * 64);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC_F16_N64, k > 0 ? 1u : 0u);
                    umma_commit(bar_empty + 8 * s);       // stage free once the MMAs have read it
                    umma_commit(bar_accf + 8 * tb);       // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: TMEM -> message -> destination row of the group's accumulator plane =====================
        const int q = warp & 3, grp = warp >> 2;          // TMEM lanes 32q .. 32q+31 = slots 32q .. 32q+31 of the block
        const float bias_c = __ldg(A.bias + lane);
        const int rot = lane & 7;
        const uint32_t my_acc = acc_base + (uint32_t)grp * (uint32_t)A.rt * 128u;
        double s1 = 0.0, s2 = 0.0;
        int g0 = 0;                                       // block counter of the CTA at the start of the tile
        bool ok = true;
        for (int tile = blockIdx.x; ok && tile < A.n_tiles; tile += gridDim.x) {
            const int b0 = __ldg(A.bptr + tile), b1 = __ldg(A.bptr + tile + 1);
            const int node0 = tile * A.rt;
            const int first = b0 + ((grp - g0) & 1);      // this group's blocks: (g0 + blk - b0) & 1 == grp
            int dst = 0xFFFF, type = 0;
            if (first < b1) { dst = __ldg(A.tdst + (size_t)first * TBS + 32 * q + lane); type = __ldg(A.btype + first); }
            for (int blk = first; blk < b1; blk += 2) {
                const int g = g0 + (blk - b0);
                int ndst = 0xFFFF, ntype = 0;
                if (blk + 2 < b1) { ndst = __ldg(A.tdst + (size_t)(blk + 2) * TBS + 32 * q + lane); ntype = __ldg(A.btype + blk + 2); }
                const bool root = type == A.n_types, live = dst != 0xFFFF;
                // mean aggregation: every edge message is scaled by 1 / deg(dst) as it is added (two planes and no order between
                // the groups: a final scaling pass would have to know which part of a row is the root term)
                float inv = 1.f;
                if (!root && live) inv = __ldg(A.inv_deg + node0 + dst);
                // the destination row's eight float4 pieces, in this lane's rotated order, BEFORE the wait for the MMAs: step
                // st touches piece (st + lane) & 7 -- the eight lanes of a quarter-warp are on eight different bank groups
                // whatever their rows are.  (No other warp touches this row: one class per quarter, one plane per group.)
                const uint32_t row = my_acc + (uint32_t)(live ? dst : 0) * 128u;
                __syncwarp();                                                 // the previous block's stores of OTHER lanes to this row
                float4 a[8];
#pragma unroll
                for (int st = 0; st < 8; ++st) a[st] = lds128f(row + (uint32_t)(((st + rot) & 7) << 4));
                const int tb = g % NT;
                if (!TGNN_TIMED(w0, mbar_wait(bar_accf + 8 * tb, (uint32_t)((g / NT) & 1)))) { timeout_flag = 1; ok = false; break; }
                tc_fence_after();
                float m[32], sm[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(tb * 64);
                tmem_ld32_issue(taddr, m);
                tmem_ld32_issue(taddr + 32u, sm);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acce + 8 * tb);                // the MMA warp may overwrite the buffer
#pragma unroll
                for (int i = 0; i < 32; ++i) m[i] = fmaf(sm[i], LO_INV, m[i]);
                // rotate left by 4 * (lane & 7) channels: u[c] = m[(c + 4 rot) & 31]
#pragma unroll
                for (int st = 0; st < 3; ++st) {
                    const bool on = (rot >> st) & 1;
                    const int sh = 4 << st;
                    float u[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) u[i] = on ? m[(i + sh) & 31] : m[i];
#pragma unroll
                    for (int i = 0; i < 32; ++i) m[i] = u[i];
                }
                if (live) {
#pragma unroll
                    for (int st = 0; st < 8; ++st) {
                        a[st].x = fmaf(m[4 * st], inv, a[st].x); a[st].y = fmaf(m[4 * st + 1], inv, a[st].y);
                        a[st].z = fmaf(m[4 * st + 2], inv, a[st].z); a[st].w = fmaf(m[4 * st + 3], inv, a[st].w);
                    }
#pragma unroll
                    for (int st = 0; st < 8; ++st) sts128f(row + (uint32_t)(((st + rot) & 7) << 4), a[st]);
                }
                dst = ndst; type = ntype;
            }
            if (!ok) break;
            g0 += b1 - b0;
            const long long t_fin = clock64();
            named_bar_sync(1, EPI_WARPS * 32);                                // every message of the tile is in
            finish_tile(A, acc_base, tile, warp, lane, bias_c, s1, s2);
            named_bar_sync(1, EPI_WARPS * 32);                                // zeroed before the next tile's first add
            w1 += clock64() - t_fin;
        }
        if (A.part) {
            A.part[((size_t)blockIdx.x * EPI_WARPS + warp) * 64 + lane] = s1;
            A.part[((size_t)blockIdx.x * EPI_WARPS + warp) * 64 + 32 + lane] = s2;
        }
    }
    if (A.dbg && blockIdx.x == 0 && lane == 0) {
        long long* d = A.dbg + warp * 4;
        d[0] = clock64() - t_start; d[1] = w0; d[2] = w1; d[3] = w2;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NT * 64));
    }
    if (timeout_flag && tid == 0 && A.err) { *reinterpret_cast<volatile int*>(A.err) = TGNN_DEVERR_PIPELINE; __threadfence_system(); }
}

// The same layer on the same blocks with fp32 FMAs on CUDA cores (no range limit): taken only when a range flag is
// raised.  Thread = slot; a block's 128 destinations are distinct, so plain read-modify-write per block is race free.
__global__ void __launch_bounds__(TBS)
k_conv_t_wide(ConvTArgs A) {
    if (!((A.flag_x && *A.flag_x) || (A.flag_w && *A.flag_w))) return;
    extern __shared__ __align__(16) float acc[];          // [rt][WS]: odd row stride, rows of different threads on different banks
    constexpr int WS = 33;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float bias_c = __ldg(A.bias + lane);
    double s1 = 0.0, s2 = 0.0;
    for (int i = tid; i < A.rt * WS; i += TBS) acc[i] = 0.f;
    __syncthreads();
    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        const int b0 = A.bptr[tile], b1 = A.bptr[tile + 1], node0 = tile * A.rt;
        for (int blk = b0; blk < b1; ++blk) {
            const int src = A.tsrc[(size_t)blk * TBS + tid], dst = A.tdst[(size_t)blk * TBS + tid], type = A.btype[blk];
            if (src >= 0) {
                const float* x = A.xin + (size_t)src * F;
                const float* W = A.tab32 + (size_t)type * (F * F);
                float m[32];
#pragma unroll
                for (int n = 0; n < 32; ++n) m[n] = 0.f;
                for (int k = 0; k < 32; ++k) {
                    const float xv = __ldg(x + k);
#pragma unroll
                    for (int n = 0; n < 32; ++n) m[n] = fmaf(xv, __ldg(W + k * 32 + n), m[n]);
                }
                float* row = acc + dst * WS;
                const float inv = type == A.n_types ? 1.f : __ldg(A.inv_deg + node0 + dst);
#pragma unroll
                for (int n = 0; n < 32; ++n) row[n] = fmaf(m[n], inv, row[n]);
            }
            __syncthreads();
        }
        const int rows = min(A.rt, A.n_own - node0);
        for (int r = warp; r < rows; r += TBS / 32) {
            float v = leaky(acc[r * WS + lane] + bias_c);
            if (!row_kept(A.mask, node0 + r)) v = 0.f;
            acc[r * WS + lane] = 0.f;
            A.out[(size_t)(node0 + r) * F + lane] = v;
            s1 += (double)v;
            s2 += (double)v * (double)v;
        }
        __syncthreads();
    }
    if (A.part) {                                     // same partial layout as k_conv_t (8 rows per CTA; this kernel fills 4)
        A.part[((size_t)blockIdx.x * EPI_WARPS + warp) * 64 + lane] = s1;
        A.part[((size_t)blockIdx.x * EPI_WARPS + warp) * 64 + 32 + lane] = s2;
        A.part[((size_t)blockIdx.x * EPI_WARPS + 4 + warp) * 64 + lane] = 0.0;
        A.part[((size_t)blockIdx.x * EPI_WARPS + 4 + warp) * 64 + 32 + lane] = 0.0;
    }
}

size_t conv_t_smem(int rt) { return (size_t)EPI_GROUPS * rt * 128 + (size_t)NS * STAGE_BYTES + 1024; }

}  // namespace

int conv_t_blocks(int t_tiles, int sm_count) { return t_tiles < sm_count ? (t_tiles < 1 ? 1 : t_tiles) : sm_count; }

void launch_conv_t(const ConvArgs& c, const Graph& g, const uint32_t* tabT, const float* tab32, int* err, int sm_count, cudaStream_t st,
                   long long* dbg) {
    static PerDeviceOnce once;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_t, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_t_smem(512)));
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_t_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 33 * 4));
    });
    ConvTArgs a{};
    a.xh = c.xh; a.xin = c.xin; a.tabT = tabT; a.tab32 = tab32; a.n_types = c.n_types;
    a.bptr = g.t_bptr.as<int>(); a.btype = g.t_btype.as<int>(); a.tsrc = g.t_src.as<int>(); a.tdst = g.t_dst.as<unsigned short>();
    a.inv_deg = c.inv_deg; a.bias = c.bias; a.out = c.out; a.part = c.part;
    a.flag_x = c.flag_x; a.flag_w = c.flag_w; a.err = err; a.mask = c.mask; a.dbg = dbg;
    a.n_own = c.n_own; a.n_tiles = g.t_tiles; a.rt = g.t_rows;
    const int blocks = conv_t_blocks(g.t_tiles, sm_count);
    k_conv_t<<<blocks, CT_THREADS, conv_t_smem(g.t_rows), st>>>(a);
    TGNN_CUDA(cudaGetLastError());
    // stand-by for the range guard: same grid (same BatchNorm partial layout), exits at once unless a flag is raised
    k_conv_t_wide<<<blocks, TBS, (size_t)g.t_rows * 33 * 4, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
