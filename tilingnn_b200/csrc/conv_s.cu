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
// Persistent CTAs (one per SM) walk tiles of 128 destinations; 26 warps, warp specialised:
//   loaders (2 warps): per pass one TMA bulk copy (cp.async.bulk) for the pre-swizzled weight tiles, one for the
//     row-offset table (and the root pass's contiguous rows), 16-byte cp.async for the scattered source rows, into
//     a 640-row ring / 4-slot ring; completion = mbarrier transaction bytes + cp.async.mbarrier.arrive, so 4 passes
//     of gathers are in flight without holding registers.  (Per-row bulk copies were measured 2.6x slower: the
//     instruction is uniform-datapath, 32 lanes with different addresses serialise.)
//   transformers (2 groups x 8 warps, group g = passes with parity g = A stage g): per (row, 16-byte chunk) sum the
//     row's sources from the ring, hi/lo TF32 split, swizzled store into the A operand stage, fence, arrive
//   MMA warp: per pass 12 x tcgen05.mma.kind::tf32 (M=128, N=32; 3xTF32), commits free the stage and the slot
//   the root term x_i root is one more pass into a second accumulator (columns 32..63); TMEM is double buffered
//   epilogue (4 warps): tcgen05.ld, * 1/deg + root + bias, LeakyReLU, store, BatchNorm partial sums (fp64),
//     overlapping the next tile's passes
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "layouts.cuh"
#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {
using namespace tc;

constexpr int SA_TILE = 16384;                 // bytes of one A tile (hi or lo): 128 rows x 128 B
constexpr int SB_TILE = 4096;                  // bytes of one B tile (hi or lo): 32 rows x 128 B
constexpr int D_SLOTS = 4;                     // passes in flight (ring of weight tiles / offset rows / barriers)
constexpr int RING_ROWS = 640;                 // gathered source rows in flight (80 KB): 4 passes of <= 160 rows
constexpr int N_LOAD = 4;                      // row-gather loader warps (cooperate on every pass)
constexpr int W_BULK = N_LOAD;                 // one warp (one lane) issuing the TMA bulk copies
constexpr int W_X0 = W_BULK + 1;               // first transformer warp
constexpr int N_XGRP = 8, N_XFORM = 2 * N_XGRP;   // two transformer groups of 8 warps (one per A stage)
constexpr int W_MMA = W_X0 + N_XFORM, W_EPI0 = W_MMA + 1;
constexpr int CS_THREADS = (W_EPI0 + 4) * 32;  // 26 warps
constexpr int OFF_A = 0;                                        // [2 stages][hi|lo]
constexpr int OFF_B = OFF_A + 4 * SA_TILE;                      // [D_SLOTS][hi|lo]
constexpr int OFF_RING = OFF_B + D_SLOTS * 2 * SB_TILE;         // [RING_ROWS][128 B]
constexpr int OFF_OFFB = OFF_RING + RING_ROWS * 128;            // [D_SLOTS][S_OFF_STRIDE] uint16
constexpr int OFF_META = OFF_OFFB + D_SLOTS * S_OFF_STRIDE * 2; // [D_SLOTS] int4 {first ring row of the pass, root flag, -, -}
constexpr int OFF_EPI = OFF_META + D_SLOTS * 16;                // scratch [4][32*33] float, red [4][2][32] double
constexpr int CS_SMEM = OFF_EPI + 4 * 32 * 33 * 4 + 4 * 2 * 32 * 8 + 1024;

__device__ __forceinline__ float4 ld_row4(const float* base, int row, int q) {
    return __ldg(reinterpret_cast<const float4*>(base + (size_t)row * F) + q);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

struct ConvSArgs {
    const float* xin;            // [n_rows][32]
    const float* tabS;           // [K+1][hi|lo] swizzled smem images of W_t^T [32 n][32 k]  (entry K = root^T)
    int n_types;
    const int* pptr; const int* ptype; const int* pbase; const unsigned short* off; const int* ssrc;
    const float* inv_deg; const float* bias;
    float* out; double* part; int* error_flag;
    const uint8_t* mask;
    long long* dbg;              // optional per-role wait/total cycle counters of CTA 0 (TGNN_CONVS_DBG=1)
    int n_own, n_tiles, d_eff;   // d_eff: passes in flight such that d_eff * (longest pass) <= RING_ROWS
};

#ifdef TGNN_CONVS_TIMING
#define TIMED(acc, expr) ([&]() { const long long _t = clock64(); const bool _r = (expr); (acc) += clock64() - _t; return _r; })()
#else
#define TIMED(acc, expr) (expr)
#endif

__global__ void __launch_bounds__(CS_THREADS, 1)
k_conv_s(ConvSArgs A) {
    long long w0 = 0, w1 = 0, w2 = 0;                  // cycles spent in this role's barrier waits
    const long long t_start = clock64();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * D_SLOTS + 4 + 4];   // raw_full[D], pass_done[D], a_full[2], -, acc_full[2], acc_empty[2]
    __shared__ uint32_t tmem_base_smem;
    __shared__ int timeout_flag;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_rf = smem_u32(&bars[0]), bar_re = smem_u32(&bars[D_SLOTS]);
    const uint32_t bar_af = smem_u32(&bars[2 * D_SLOTS]);
    const uint32_t bar_cf = smem_u32(&bars[2 * D_SLOTS + 4]), bar_ce = smem_u32(&bars[2 * D_SLOTS + 6]);
    constexpr int D = D_SLOTS;

    if (tid == 0) {
        for (int i = 0; i < D_SLOTS; ++i) { mbar_init(bar_rf + 8 * i, N_LOAD * 32 + 1); mbar_init(bar_re + 8 * i, N_XGRP + 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_af + 8 * i, N_XGRP); mbar_init(bar_cf + 8 * i, 1); mbar_init(bar_ce + 8 * i, 4); }
        timeout_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < N_LOAD) {
        // ===================== loaders: the 4 warps cooperate on EVERY pass =====================
        // The number of cp.async a thread can have outstanding is small, so a pass's ~640 16-byte row requests are
        // spread over all 128 loader threads (<= 10 per lane): instruction k of warp w covers edges 16k + 4w + (lane/8),
        // 8 lanes per 128-byte row.  Each lane fetches the source indices of ITS edges one pass ahead (two register
        // buffers, explicit branch paths so the loads do not feed a select).
        const int c = lane & 7, sub = lane >> 3;
        const uint32_t ring_c = smem_base + OFF_RING + c * 16;
        const float* xin_c = A.xin + 4 * c;
        const int e_lane = 4 * warp + sub;                 // edge handled by this lane in instruction k: 16k + e_lane
        int s = 0, rows = 0;                               // pass counter / ring rows consumed so far
        int idxA[10], idxB[10];
        bool par = false;                                  // pending pass's indices live in (par ? idxA : idxB)
        int pd_tile = -1, pd_off = 0, pd_len = 0, pd_type = 0, pd_s = 0, pd_rows = 0;
        auto issue = [&]() -> bool {
            const bool root = pd_len < 0;
            const int slot = pd_s & (D - 1);
            const uint32_t bar = bar_rf + 8 * slot;
            if (!TIMED(w0, mbar_wait_relaxed(bar_re + 8 * slot, (uint32_t)(((pd_s / D) & 1) ^ 1)))) return false;
            const int ring0 = pd_rows % RING_ROWS;
            int len = pd_len;
            if (root) { len = A.n_own - pd_tile * S_BM; len = len > S_BM ? S_BM : len; }
            if (!root) {
                if (par) {
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        const int e = 16 * k + e_lane;
                        if (e < len) { int rr = ring0 + e; if (rr >= RING_ROWS) rr -= RING_ROWS;
                                       cp_async16(ring_c + (uint32_t)rr * 128, xin_c + (size_t)idxA[k] * F); }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        const int e = 16 * k + e_lane;
                        if (e < len) { int rr = ring0 + e; if (rr >= RING_ROWS) rr -= RING_ROWS;
                                       cp_async16(ring_c + (uint32_t)rr * 128, xin_c + (size_t)idxB[k] * F); }
                    }
                }
            }
            // asynchronous arrive: counts once all cp.async issued by this lane have landed
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
            return true;
        };
        bool ok = true;
        for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x) {
            const int p0 = __ldg(A.pptr + tile), np = __ldg(A.pptr + tile + 1) - p0;
            int pb[4], pt[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                pb[j] = (32 * j + lane <= np) ? __ldg(A.pbase + p0 + 32 * j + lane) : 0;
                pt[j] = (32 * j + lane < np) ? __ldg(A.ptype + p0 + 32 * j + lane) : 0;
            }
            auto lookup = [&](const int (&arr)[4], int i) {
                const int v0 = __shfl_sync(0xffffffffu, arr[0], i & 31), v1 = __shfl_sync(0xffffffffu, arr[1], i & 31);
                const int v2 = __shfl_sync(0xffffffffu, arr[2], i & 31), v3 = __shfl_sync(0xffffffffu, arr[3], i & 31);
                const int j = i >> 5;
                return j == 0 ? v0 : (j == 1 ? v1 : (j == 2 ? v2 : v3));
            };
            int base = lookup(pb, 0);
            for (int q = 0; q <= np && ok; ++q, ++s) {
                const bool root = q == np;
                const int nbase = root ? 0 : lookup(pb, q + 1);
                const int len = root ? S_BM : nbase - base;
                const int type = root ? A.n_types : lookup(pt, q);
                if (!root) {                                // request this pass's indices into the free buffer
                    const int* sp = A.ssrc + base + e_lane;
                    if (par) {
#pragma unroll
                        for (int k = 0; k < 10; ++k) if (16 * k + e_lane < len) idxB[k] = __ldg(sp + 16 * k);
                    } else {
#pragma unroll
                        for (int k = 0; k < 10; ++k) if (16 * k + e_lane < len) idxA[k] = __ldg(sp + 16 * k);
                    }
                }
                if (pd_tile >= 0) ok = issue();             // issue the pass requested one pass ago
                pd_tile = tile; pd_off = p0 + q; pd_len = root ? -1 : len; pd_type = type; pd_s = s; pd_rows = rows;
                par = !par;
                rows += len;
                base = nbase;
            }
        }
        if (ok && pd_tile >= 0) ok = issue();
        if (!ok) timeout_flag = 1;
    } else if (warp == W_BULK) {
        // ===================== bulk loader: the contiguous pieces of every pass, one TMA bulk copy each ==========
        // pre-swizzled weight image (8 KB), row-offset table (272 B), and for the root pass the tile's own rows;
        // completion is counted in mbarrier transaction bytes.  (cp.async.bulk is a uniform-datapath instruction:
        // one lane issues it, which is why the scattered rows are NOT fetched this way.)
        int s = 0, rows = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x) {
            const int p0 = __ldg(A.pptr + tile), np = __ldg(A.pptr + tile + 1) - p0;
            int pb[4], pt[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                pb[j] = (32 * j + lane <= np) ? __ldg(A.pbase + p0 + 32 * j + lane) : 0;
                pt[j] = (32 * j + lane < np) ? __ldg(A.ptype + p0 + 32 * j + lane) : 0;
            }
            auto lookup = [&](const int (&arr)[4], int i) {
                const int v0 = __shfl_sync(0xffffffffu, arr[0], i & 31), v1 = __shfl_sync(0xffffffffu, arr[1], i & 31);
                const int v2 = __shfl_sync(0xffffffffu, arr[2], i & 31), v3 = __shfl_sync(0xffffffffu, arr[3], i & 31);
                const int j = i >> 5;
                return j == 0 ? v0 : (j == 1 ? v1 : (j == 2 ? v2 : v3));
            };
            int base = lookup(pb, 0);
            for (int q = 0; q <= np && ok; ++q, ++s) {
                const bool root = q == np;
                const int nbase = root ? 0 : lookup(pb, q + 1);
                const int type = root ? A.n_types : lookup(pt, q);
                int len = root ? A.n_own - tile * S_BM : nbase - base;
                if (root && len > S_BM) len = S_BM;
                const int slot = s & (D - 1);
                const uint32_t bar = bar_rf + 8 * slot;
                if (!TIMED(w0, mbar_wait_relaxed(bar_re + 8 * slot, (uint32_t)(((s / D) & 1) ^ 1)))) { ok = false; break; }
                const int ring0 = rows % RING_ROWS;
                if (lane == 0) {
                    sts128i(smem_base + OFF_META + slot * 16, make_int4(ring0, root ? 1 : 0, 0, 0));
                    mbar_arrive_expect_tx(bar, (uint32_t)(2 * SB_TILE + (root ? len * 128 : S_OFF_STRIDE * 2)));
                    bulk_g2s(smem_base + OFF_B + (uint32_t)slot * 2 * SB_TILE, A.tabS + (size_t)type * 2048, 2 * SB_TILE, bar);
                    if (!root) {
                        bulk_g2s(smem_base + OFF_OFFB + (uint32_t)slot * S_OFF_STRIDE * 2, A.off + (size_t)(p0 + q) * S_OFF_STRIDE,
                                 S_OFF_STRIDE * 2, bar);
                    } else {
                        const int first = min(len, RING_ROWS - ring0);
                        const float* src = A.xin + (size_t)tile * S_BM * F;
                        bulk_g2s(smem_base + OFF_RING + (uint32_t)ring0 * 128, src, first * 128, bar);
                        if (len > first) bulk_g2s(smem_base + OFF_RING, src + (size_t)first * F, (len - first) * 128, bar);
                    }
                }
                __syncwarp();
                rows += root ? S_BM : len;
                base = nbase;
            }
        }
        if (!ok) timeout_flag = 1;
    } else if (warp < W_MMA) {
        // ===================== transformers: group xg owns A stage xg and the passes with (s & 1) == xg ==========
        // item (row = tt/8 + 32 j, chunk = tt%8), tt = thread index inside the group
        const int xg = (warp - W_X0) / N_XGRP;
        const int tt = tid - (W_X0 + xg * N_XGRP) * 32, c = tt & 7, rbase = tt >> 3;
        const uint32_t item_off = sw128_off(rbase, c);          // + 4096 j for row rbase + 32 j (same row & 7)
        const uint32_t ring_c = smem_base + OFF_RING + c * 16;
        const uint32_t sa_hi = smem_base + OFF_A + (xg * 2) * SA_TILE + item_off, sa_lo = sa_hi + SA_TILE;
        uint32_t dirty = 0xFu;                                  // items whose last stored value was non-zero
        int s = 0;
        bool ok = true;
        int np_next = (int)blockIdx.x < A.n_tiles ? __ldg(A.pptr + blockIdx.x + 1) - __ldg(A.pptr + blockIdx.x) : 0;
        for (int tile = blockIdx.x; tile < A.n_tiles && ok; tile += gridDim.x) {
            const int np = np_next;
            const int ntile = tile + gridDim.x;
            if (ntile < A.n_tiles) np_next = __ldg(A.pptr + ntile + 1) - __ldg(A.pptr + ntile);
            for (int q = 0; q <= np; ++q, ++s) {
                if ((s & 1) != xg) continue;
                const int slot = s & (D - 1);
                if (!TIMED(w0, mbar_wait(bar_rf + 8 * slot, (uint32_t)((s / D) & 1)))) { ok = false; break; }
                const int ring0 = lds128i(smem_base + OFF_META + slot * 16).x;
                const bool root = q == np;
                float4 v[4];
                uint32_t nz = 0;
                if (!root) {
                    const uint32_t ob = smem_base + OFF_OFFB + (uint32_t)slot * S_OFF_STRIDE * 2 + 2 * rbase;
                    int es[4], cnt[4], mx = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        es[j] = (int)lds_u16(ob + 64 * j);
                        cnt[j] = (int)lds_u16(ob + 64 * j + 2) - es[j];
                        mx = max(mx, cnt[j]);
                        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (cnt[j] > 0) nz |= 1u << j;
                        es[j] += ring0;
                    }
                    for (int t = 0; t < mx; ++t) {              // the four items' loads are independent -> in flight together
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (t < cnt[j]) {
                                int rr = es[j] + t; if (rr >= RING_ROWS) rr -= RING_ROWS;
                                const float4 w = lds128f(ring_c + (uint32_t)rr * 128);
                                v[j].x += w.x; v[j].y += w.y; v[j].z += w.z; v[j].w += w.w;
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = rbase + 32 * j;
                        int rr = ring0 + r; if (rr >= RING_ROWS) rr -= RING_ROWS;
                        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (tile * S_BM + r < A.n_own) { v[j] = lds128f(ring_c + (uint32_t)rr * 128); nz |= 1u << j; }
                    }
                }
                // A stage xg was last read by the MMAs of pass s-2: wait for that pass's "done" barrier
                if (s >= 2 && !TIMED(w1, mbar_wait(bar_re + 8 * ((s - 2) & (D - 1)), (uint32_t)(((s - 2) / D) & 1)))) { ok = false; break; }
                const uint32_t need = nz | dirty;               // rows that are and stay zero need no store at all
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (need & (1u << j)) {
                        // truncation split: hi keeps the top 19 bits (what kind::tf32 reads), lo = x - hi is exact and
                        // the tensor core truncates it to TF32 itself (relative residual ~2^-21)
                        uint4 hi, lo;
                        hi.x = __float_as_uint(v[j].x) & 0xFFFFE000u; lo.x = __float_as_uint(v[j].x - __uint_as_float(hi.x));
                        hi.y = __float_as_uint(v[j].y) & 0xFFFFE000u; lo.y = __float_as_uint(v[j].y - __uint_as_float(hi.y));
                        hi.z = __float_as_uint(v[j].z) & 0xFFFFE000u; lo.z = __float_as_uint(v[j].z - __uint_as_float(hi.z));
                        hi.w = __float_as_uint(v[j].w) & 0xFFFFE000u; lo.w = __float_as_uint(v[j].w - __uint_as_float(hi.w));
                        sts128(sa_hi + 4096 * j, hi);
                        sts128(sa_lo + 4096 * j, lo);
                    }
                }
                dirty = nz;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_af + 8 * xg); mbar_arrive(bar_re + 8 * slot); }
            }
        }
        if (!ok) timeout_flag = 1;
    } else if (warp == W_MMA) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t IDESC = umma_idesc_tf32(32);
            // One thread issues every MMA of the CTA, so its scalar instruction stream bounds the pass rate: all
            // descriptors are "base + small integer" (the address field is bits 0..13 in 16-byte units and every
            // operand lives below 256 KB, so plain 64-bit adds never carry out of the field).
            const uint64_t dA = umma_desc_sw128(smem_base + OFF_A), dB = umma_desc_sw128(smem_base + OFF_B);
            int s = 0, it = 0, q = 0;
            for (int tile = blockIdx.x; tile < A.n_tiles;) {
                const int slot = s & (D - 1), st = s & 1, ab = it & 1;
                if (q == 0 && !TIMED(w2, mbar_wait(bar_ce + 8 * ab, (uint32_t)(((it >> 1) & 1) ^ 1)))) { timeout_flag = 1; break; }
                if (!TIMED(w0, mbar_wait(bar_rf + 8 * slot, (uint32_t)((s / D) & 1)))) { timeout_flag = 1; break; }
                const bool root = lds128i(smem_base + OFF_META + slot * 16).y != 0;
                if (!TIMED(w1, mbar_wait(bar_af + 8 * st, (uint32_t)((s >> 1) & 1)))) { timeout_flag = 1; break; }
                fence_proxy_async();                   // weight tiles were written by the TMA engine / loaders' cp.async
                tc_fence_after();
                const uint64_t dah = dA + (uint64_t)(st * (2 * SA_TILE / 16)), dal = dah + SA_TILE / 16;
                const uint64_t dbh = dB + (uint64_t)(slot * (2 * SB_TILE / 16)), dbl = dbh + SB_TILE / 16;
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * 64) + (root ? 32u : 0u);
                if (root || q == 0) umma_tf32_ovw(tmem_d, dal, dbh, IDESC); else umma_tf32_acc(tmem_d, dal, dbh, IDESC);
                umma_tf32_acc(tmem_d, dah, dbl, IDESC);
                umma_tf32_acc(tmem_d, dah, dbh, IDESC);
#pragma unroll
                for (int ks = 1; ks < 4; ++ks) {       // 8 tf32 = 32 bytes = 2 descriptor units along K inside the swizzle atom
                    umma_tf32_acc(tmem_d, dal + 2 * ks, dbh + 2 * ks, IDESC);
                    umma_tf32_acc(tmem_d, dah + 2 * ks, dbl + 2 * ks, IDESC);
                    umma_tf32_acc(tmem_d, dah + 2 * ks, dbh + 2 * ks, IDESC);
                }
                umma_commit(bar_re + 8 * slot);        // pass done: operand stage, weight tile and ring rows are free
                ++s; ++q;
                if (root) { umma_commit(bar_cf + 8 * ab); tile += gridDim.x; ++it; q = 0; }
            }
        }
    } else {
        // ===================== epilogue warps: TMEM lane quarter q4 = warp % 4 =====================
        const int q4 = warp & 3, etid = (warp - W_EPI0) * 32 + lane;
        const uint32_t sc = smem_base + OFF_EPI + (uint32_t)q4 * (32 * 33 * 4);
        const uint32_t red = smem_base + OFF_EPI + 4 * 32 * 33 * 4;
        int it = 0;
        for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const int np = __ldg(A.pptr + tile + 1) - __ldg(A.pptr + tile);
            if (!TIMED(w0, mbar_wait_relaxed(bar_cf + 8 * ab, (uint32_t)((it >> 1) & 1)))) { timeout_flag = 1; break; }
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
                for (int r = 0; r < nv; ++r) { const double x = (double)lds_f32(sc + 4 * (r * 33 + lane)); s1 += x; s2 += x * x; }
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
    }
}

}  // namespace

void launch_conv_s(const ConvArgs& c, const Graph& g, const float* tabS, int* error_flag, int sm_count, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_conv_s, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
    });
    ConvSArgs a{};
    a.xin = c.xin; a.tabS = tabS; a.n_types = g.n_types;
    a.pptr = g.s_pptr.as<int>(); a.ptype = g.s_ptype.as<int>(); a.pbase = g.s_pbase.as<int>();
    a.off = g.s_off.as<unsigned short>(); a.ssrc = g.s_src.as<int>();
    a.inv_deg = c.inv_deg; a.bias = c.bias; a.out = c.out; a.part = c.part; a.error_flag = error_flag; a.mask = c.mask;
    a.n_own = c.n_own; a.n_tiles = g.s_tiles;
    a.d_eff = D_SLOTS;
    static long long* dbg = nullptr;
    static const bool want_dbg = getenv("TGNN_CONVS_DBG") != nullptr;
    if (want_dbg && !dbg) TGNN_CUDA(cudaMalloc(&dbg, 32 * 4 * sizeof(long long)));
    a.dbg = dbg;
    TGNN_CHECK(g.s_max_pass * D_SLOTS <= RING_ROWS, "conv_s: pass too long for the shared-memory ring");
    k_conv_s<<<std::min(g.s_tiles, sm_count), CS_THREADS, CS_SMEM, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
    if (want_dbg) {
        static int calls = 0;
        if (++calls == 30) {                           // a warmed-up launch
            long long h[32 * 4];
            TGNN_CUDA(cudaStreamSynchronize(st));
            TGNN_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
            const char* role[] = {"loader", "xform", "mma", "epilogue"};
            for (int w = 0; w < CS_THREADS / 32; ++w) {
                const int r = w < N_LOAD ? 0 : (w < W_MMA ? 1 : (w == W_MMA ? 2 : 3));
                fprintf(stderr, "conv_s dbg warp %2d %-8s total %9lld  wait0 %9lld  wait1 %9lld  wait2 %9lld\n", w, role[r], h[4 * w],
                        h[4 * w + 1], h[4 * w + 2], h[4 * w + 3]);
            }
        }
    }
}

}  // namespace tgnn
