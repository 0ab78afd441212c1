// Dense stages of the final MLP (TilinGNN.py:45-46,74-76 of the reference) on the 5th-generation
// tensor cores:  out = LeakyReLU( BN_in(A) @ W^T + b )  and per-column BatchNorm partial sums.
//
// Persistent CTAs (one per SM) walk tiles of 128 rows x all N_out columns.  The fp32 accumulator lives in
// TMEM (128 lanes x N_out columns), double buffered so the epilogue of one tile overlaps the next main loop.  K is walked in slabs of 32 (= one 128-byte SWIZZLE_128B atom row):
//   producers (warps 0-7): A slab  global -> registers -> lazy BatchNorm -> hi/lo TF32 split -> swizzled smem
//   bulk warp (warp 8, one lane): W slab = ONE TMA bulk copy of the pre-swizzled hi|lo image (mbarrier tx bytes)
//   MMA issuer (warp 8, one lane): 12 x tcgen05.mma.kind::tf32 per slab (4 K-steps of 8 x {lo*hi, hi*lo, hi*hi}),
//                          tcgen05.commit -> mbarrier frees the smem stage / publishes the accumulator
//   epilogue (warps 9-12): tcgen05.ld 32 columns at a time -> bias, LeakyReLU -> global, column sums in fp64.
// 3xTF32 keeps fp32-level accuracy (single-pass TF32 would break the 1e-4 parity bar).
//
// k_dense_tc<N, true> is the same pipeline on fp16 two-term splits (kind::f16): x = hi + lo with hi = fp16(x),
// lo = fp16(x - hi) (unscaled: the inputs are BatchNorm outputs / layer outputs of O(1), so lo's subnormal spacing 2^-24
// is an ABSOLUTE error far below fp32 rounding of the sums), weights scaled by 2^6 at pack time so their lo parts stay in
// the normal range (undone on the accumulator, exact).  A slab's A tile is ONE 128-byte-row tile {hi(32) | lo(32)} and so
// is the weight tile: half the shared-memory bytes, and 6 MMAs (K = 16) per slab instead of 12 (K = 8) -- the tensor
// time per tile halves (a tcgen05.mma costs max(44.6, N/2) cycles whatever the kind).  Values outside the fp16 range
// (checked by the producers, |x| > 60000) raise a flag; the 3xTF32 kernel, launched right behind, then redoes the stage
// (it exits at once otherwise).  Weights outside the range (checked at pack time) keep the stage on 3xTF32.
// Every mbarrier wait is bounded: on a timeout the kernel raises a device-side error flag and exits instead of
// hanging the GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <cuda_fp16.h>

#include "bn_fin.cuh"
#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {

constexpr int BM = 128;           // rows per tile (UMMA M)
constexpr int BK = 32;            // K per slab (128 bytes of tf32)
constexpr int A_TILE_BYTES = BM * BK * 4;            // 16 KB
constexpr int N_PROD_WARPS = 8, BULK_WARP = 8, MMA_WARP = 9, EPI_WARP0 = 10;
constexpr int N_EPI_WARPS = 4;
constexpr int NTHREADS = (EPI_WARP0 + N_EPI_WARPS) * 32;   // 8 A-producer warps, 1 bulk (TMA) warp, 1 MMA warp, 4 epilogue warps
constexpr int SCRATCH_FLOATS = 32 * 36;              // per epilogue warp: one 32x32 block, row stride 36 (16-B aligned, conflict-free)

using namespace tc;

struct DenseTcArgs {
    const float* const* slabs;  // virtual concat: K/32 slab pointers [n_rows][32]
    const float* a;             // else plain [n][K]
    int virtual_concat;
    const float* in_coef;       // [4][K] lazy BatchNorm of the input, or nullptr
    const float* w_img;         // [K/32 slabs][hi|lo][N_out x 128 B] SWIZZLE_128B images of the TF32-split weights
    const float* bias;          // [N_out]
    float* out;                 // [n][N_out]
    double* part;               // [gridDim.x][2][N_out] or nullptr
    int* error_flag;
    long long* dbg;             // optional per-role wait counters of CTA 0 (TGNN_DENSE_DBG=1)
    int* range_flag;            // fp16 variant: raised when an input value is outside the fp16 range; tf32 variant: run only if raised ...
    int standby;                // ... when standby != 0 (else the tf32 variant always runs)
    int n, K;
    const uint8_t* mask;        // node mask or null: masked rows are written as 0 (and add nothing to the column sums)
    BnFin fin;                  // fin.part != null: finish the input's BatchNorm (C = K) in the prologue, publish it to in_coef
};

constexpr float H_WSCALE = 64.f, H_WSCALE_INV = 1.f / 64.f;   // fp16 variant: weights are packed as w * 2^6 (their lo parts stay normal)
template <int NOUT, bool HALF> struct DenseCfg {
    // tf32: A = hi tile + lo tile (16 KB each), B = hi tile + lo tile (NOUT x 128 B each);  fp16: one tile each, rows {hi | lo}
    static constexpr int B_TILE_BYTES = NOUT * BK * 4;
    static constexpr int STAGE_BYTES = HALF ? A_TILE_BYTES + B_TILE_BYTES : 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = HALF ? (NOUT >= 256 ? 4 : 6) : (NOUT >= 256 ? 2 : (NOUT >= 128 ? 3 : 4));
    static constexpr int EPI_BYTES = N_EPI_WARPS * SCRATCH_FLOATS * 4 + 4 * 2 * NOUT * 4 + NOUT * 4;   // scratch, red (fp32 block sums), bias
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024;
};

// Persistent, warp-specialised:  producers (warps 0-7) -> smem ring -> MMA warp -> TMEM (double buffered)
// -> epilogue warps (9-12).  The epilogue of tile i overlaps the main loop of tile i+1.
#ifdef TGNN_DENSE_TIMING     // role wait counters (build with TGNN_NVCC_EXTRA=-DTGNN_DENSE_TIMING, run with TGNN_DENSE_DBG=1)
#define DTIMED(acc, expr) ([&]() { const long long _t = clock64(); const bool _r = (expr); (acc) += clock64() - _t; return _r; })()
#else
#define DTIMED(acc, expr) (expr)
#endif

__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
// two values -> packed fp16 pairs {hi(v0), hi(v1)} and {lo(v0), lo(v1)}, lo = fp16(v - hi) unscaled
__device__ __forceinline__ void split_h2u(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int NOUT, bool HALF>
__global__ void __launch_bounds__(NTHREADS, 1)
k_dense_tc(DenseTcArgs A) {
    if (!HALF && A.standby && !(A.range_flag && *A.range_flag)) return;      // stand-by of the fp16 variant: nothing to redo
    long long w0 = 0, w1 = 0;
    const long long t_start = clock64();
    using Cfg = DenseCfg<NOUT, HALF>;
    constexpr int STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES, B_TILE_BYTES = Cfg::B_TILE_BYTES;
    constexpr uint32_t IDESC = HALF ? ((1u << 4) | ((uint32_t)(NOUT >> 3) << 17) | ((uint32_t)(UMMA_M >> 4) << 24)) : umma_idesc_tf32(NOUT);
    constexpr int TMEM_COLS = 2 * NOUT < 32 ? 32 : 2 * NOUT;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int timeout_flag;
    __shared__ const float* slab_ptr[32];              // virtual concat: slab base pointers (no dependent global load per slab)
    __shared__ __align__(16) float coef_s[4 * 256];    // input BatchNorm coefficients when this launch finishes them itself (A.fin)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (A.virtual_concat && tid < A.K / BK && tid < 32) slab_ptr[tid] = A.slabs[tid];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t epi = smem_base + STAGES * STAGE_BYTES;
    const uint32_t scratch_all = epi;                                                // [4][32*33] float
    const uint32_t red = epi + N_EPI_WARPS * SCRATCH_FLOATS * 4;                     // [4][2][NOUT] float (32-row block sums)
    const uint32_t bias_s = epi + N_EPI_WARPS * SCRATCH_FLOATS * 4 + 4 * 2 * NOUT * 4;   // [NOUT] float
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[STAGES]);
    const uint32_t bar_accf = smem_u32(&bars[2 * STAGES]), bar_acce = smem_u32(&bars[2 * STAGES + 2]);
    const int n_slabs = A.K / BK;
    const int n_tiles = (A.n + BM - 1) / BM;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, N_PROD_WARPS + 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_accf + 8 * b, 1); mbar_init(bar_acce + 8 * b, N_EPI_WARPS); }
        timeout_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < NOUT; i += NTHREADS) sts_f32(bias_s + 4 * i, __ldg(A.bias + i));
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // small graphs: the previous stage's BatchNorm is finished here (every CTA the same fixed-order sums) instead of by a
    // k_bn_finish launch in between; the stage ring is idle until the producers start, so it serves as the scratch space
    const float* in_coef = A.in_coef;
    if (A.fin.part) {
        bn_finish_block(A.fin, A.K, coef_s, reinterpret_cast<double*>(smem));
        in_coef = coef_s;
    }

    if (warp < N_PROD_WARPS) {
        // ===================== producers: 256 threads, item (row = tid/8 + 32 j, chunk = tid%8) =====================
        const int c = tid & 7, rbase = tid >> 3;
        auto load_a = [&](int tile, int s, float4 (&v)[4]) {
            const int row0 = tile * BM;
            const float* abase; size_t lda; int koff;
            if (A.virtual_concat) { abase = slab_ptr[s]; lda = F; koff = 0; }
            else { abase = A.a; lda = (size_t)A.K; koff = s * BK; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = row0 + rbase + 32 * j;
                // zero first, then a PREDICATED load straight into the destination registers: `cond ? ldg : 0` becomes
                // load-to-temp + dependent move, which makes the "prefetch" wait for its data right here
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < A.n) v[j] = __ldg(reinterpret_cast<const float4*>(abase + (size_t)r * lda + koff) + c);
            }
        };
        // A rows are prefetched THREE slabs ahead in registers (three statically named buffers used round-robin --
        // no register copies, a MOV of a register with a load in flight would wait for it): one slab ahead (16 KB per
        // SM in flight) caps the stream at ~2.4 TB/s chip-wide, well below HBM speed.
        float4 pre0[4], pre1[4], pre2[4];
        int la_tile = blockIdx.x, la_s = 0;                 // look-ahead position (next slab to request)
        auto request = [&](float4 (&dst)[4]) {
            if (la_tile < n_tiles) {
                load_a(la_tile, la_s, dst);
                if (++la_s == n_slabs) { la_s = 0; la_tile += gridDim.x; }
            }
        };
        int g = 0, tile = blockIdx.x, s = 0;
        float4 nmh = make_float4(0.f, 0.f, 0.f, 0.f), nml = nmh, nsc = nmh, nbe = nmh;
        if (in_coef) {
            const float4* cf = reinterpret_cast<const float4*>(in_coef) + c;
            const int C4 = A.K / 4;
            nmh = cf[0]; nml = cf[C4]; nsc = cf[2 * C4]; nbe = cf[3 * C4];
        }
        bool bad = false;                                   // fp16 variant: a value outside the fp16 range was seen
        auto process = [&](float4 (&buf)[4]) -> bool {
            const int st = g % STAGES, row0 = tile * BM, k0 = s * BK;
            if (!DTIMED(w0, mbar_wait(bar_empty + 8 * st, ((g / STAGES) & 1) ^ 1))) return false;
            const uint32_t sa_hi = smem_base + st * STAGE_BYTES, sa_lo = sa_hi + A_TILE_BYTES;
            // BatchNorm coefficients of this thread's 4 columns (lazy BN of the input), fetched one slab ahead
            const float4 cmh = nmh, cml = nml, csc = nsc, cbe = nbe;
            if (in_coef) {
                const int ns = (s + 1 == n_slabs) ? 0 : s + 1;
                const float4* cf = reinterpret_cast<const float4*>(in_coef + ns * BK) + c;
                const int C4 = A.K / 4;
                nmh = cf[0]; nml = cf[C4]; nsc = cf[2 * C4]; nbe = cf[3 * C4];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = rbase + 32 * j;
                float4 v = buf[j];
                if (in_coef) {
                    v.x = fmaf((v.x - cmh.x) - cml.x, csc.x, cbe.x);
                    v.y = fmaf((v.y - cmh.y) - cml.y, csc.y, cbe.y);
                    v.z = fmaf((v.z - cmh.z) - cml.z, csc.z, cbe.z);
                    v.w = fmaf((v.w - cmh.w) - cml.w, csc.w, cbe.w);
                    if (row0 + r >= A.n) v = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (HALF) {
                    // fp16 split, row layout {hi(32 halves) | lo(32 halves)}: this thread's 4 columns are 8 bytes of each half
                    uint32_t h01, h23, l01, l23;
                    split_h2u(v.x, v.y, h01, l01);
                    split_h2u(v.z, v.w, h23, l23);
                    bad |= !(fabsf(v.x) <= TG_H_LIMIT) | !(fabsf(v.y) <= TG_H_LIMIT) | !(fabsf(v.z) <= TG_H_LIMIT) | !(fabsf(v.w) <= TG_H_LIMIT);
                    sts64(sa_hi + sw128_off(r, c >> 1) + 8u * (uint32_t)(c & 1), h01, h23);
                    sts64(sa_hi + sw128_off(r, 4 + (c >> 1)) + 8u * (uint32_t)(c & 1), l01, l23);
                    continue;
                }
                // truncation split: hi = top 19 bits (what kind::tf32 reads), lo = x - hi exact; the tensor core
                // truncates lo to TF32 itself (relative residual ~2^-21)
                uint4 hi, lo;
                hi.x = __float_as_uint(v.x) & 0xFFFFE000u; lo.x = __float_as_uint(v.x - __uint_as_float(hi.x));
                hi.y = __float_as_uint(v.y) & 0xFFFFE000u; lo.y = __float_as_uint(v.y - __uint_as_float(hi.y));
                hi.z = __float_as_uint(v.z) & 0xFFFFE000u; lo.z = __float_as_uint(v.z - __uint_as_float(hi.z));
                hi.w = __float_as_uint(v.w) & 0xFFFFE000u; lo.w = __float_as_uint(v.w - __uint_as_float(hi.w));
                const uint32_t off = sw128_off(r, c);
                sts128(sa_hi + off, hi);
                sts128(sa_lo + off, lo);
            }
            request(buf);                              // the buffer is free again: fetch the slab three positions ahead
            fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * st);
            ++g;
            if (++s == n_slabs) { s = 0; tile += gridDim.x; }
            return true;
        };
        request(pre0); request(pre1); request(pre2);
        bool ok = true;
        while (ok && tile < n_tiles) {
            ok = process(pre0);
            if (ok && tile < n_tiles) ok = process(pre1);
            if (ok && tile < n_tiles) ok = process(pre2);
        }
        if (!ok) timeout_flag = 1;
        if (HALF && bad && A.range_flag) *A.range_flag = 1;  // the 3xTF32 kernel launched behind this one redoes the stage
    } else if (warp == BULK_WARP) {
        // ===================== weight slabs: ONE TMA bulk copy per slab (hi and lo tiles are adjacent) =============
        if (lane == 0) {
            int g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int s = 0; s < n_slabs; ++s, ++g) {
                    const int st = g % STAGES;
                    if (!DTIMED(w0, mbar_wait_relaxed(bar_empty + 8 * st, ((g / STAGES) & 1) ^ 1))) { timeout_flag = 1; break; }
                    constexpr int NB = HALF ? 1 : 2, NA = HALF ? 1 : 2;      // weight / A tiles per slab
                    mbar_arrive_expect_tx(bar_full + 8 * st, NB * B_TILE_BYTES);
                    bulk_g2s(smem_base + st * STAGE_BYTES + NA * A_TILE_BYTES, A.w_img + (size_t)s * NB * NOUT * BK, NB * B_TILE_BYTES,
                             bar_full + 8 * st);
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            int g = 0, it = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
                const int ab = it & 1;
                if (!DTIMED(w1, mbar_wait(bar_acce + 8 * ab, ((it >> 1) & 1) ^ 1))) { timeout_flag = 1; break; }
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * NOUT);
                for (int s = 0; s < n_slabs; ++s, ++g) {
                    const int st = g % STAGES;
                    if (!DTIMED(w0, mbar_wait(bar_full + 8 * st, (g / STAGES) & 1))) { timeout_flag = 1; ok = false; break; }
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + st * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = a_hi + (HALF ? 1 : 2) * A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
                    if (HALF) {
                        // rows {hi | lo}: 16 halves = 32 bytes per K-step; steps 0, 1 = hi, steps 2, 3 = lo
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint64_t dah = umma_desc_sw128(a_hi + ks * 32), dal = umma_desc_sw128(a_hi + 64 + ks * 32);
                            const uint64_t dbh = umma_desc_sw128(b_hi + ks * 32), dbl = umma_desc_sw128(b_hi + 64 + ks * 32);
                            umma_f16_ss(tmem_d, dal, dbh, IDESC, (s > 0 || ks > 0) ? 1u : 0u);
                            umma_f16_ss(tmem_d, dah, dbl, IDESC, 1u);
                            umma_f16_ss(tmem_d, dah, dbh, IDESC, 1u);
                        }
                    } else
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ++ks) {
                        const uint32_t kb = ks * 32;   // 8 tf32 = 32 bytes along K inside the swizzle atom
                        const uint64_t dah = umma_desc_sw128(a_hi + kb), dal = umma_desc_sw128(a_lo + kb);
                        const uint64_t dbh = umma_desc_sw128(b_hi + kb), dbl = umma_desc_sw128(b_lo + kb);
                        umma_tf32(tmem_d, dal, dbh, IDESC, (s > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(tmem_d, dah, dbl, IDESC, 1u);
                        umma_tf32(tmem_d, dah, dbh, IDESC, 1u);
                    }
                    umma_commit(bar_empty + 8 * st);   // smem stage reusable once these MMAs have read it
                }
                umma_commit(bar_accf + 8 * ab);        // accumulator of this tile complete
            }
        }
    } else {
        // ===================== epilogue warps: TMEM lane quarter q = warp % 4, tile row = 32 q + lane =====================
        const int q = warp & 3, ew = warp - EPI_WARP0, etid = ew * 32 + lane;
        const uint32_t sc = scratch_all + (uint32_t)ew * SCRATCH_FLOATS * 4;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            if (!DTIMED(w0, mbar_wait(bar_accf + 8 * ab, (it >> 1) & 1))) { timeout_flag = 1; break; }
            tc_fence_after();
            const int row0 = tile * BM, row = row0 + 32 * q + lane;
            const bool live = row < A.n;
            int nv = A.n - (row0 + 32 * q);
            nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
            const unsigned keepbits = __ballot_sync(0xffffffffu, !live || row_kept(A.mask, row));   // bit r: row 32 q + r is kept
#pragma unroll 1
            for (int c0 = 0; c0 < NOUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(ab * NOUT + c0), v);
                // transpose through the warp's 32x36 scratch block (8 x STS.128 per lane): every global store below is ONE
                // coalesced 128-byte row segment (lane = column); bias, LeakyReLU and the column sums happen in the same
                // loop.  Sums are fp32 over the block's 32 rows, then fp64 across blocks / tiles (these BatchNorms are
                // well conditioned; the ill-conditioned conv/GIN ones keep fp64 throughout).
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts128(sc + 4 * (lane * 36 + 4 * j), make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                __syncwarp();
                const float bias_c = lds_f32(bias_s + 4 * (c0 + lane));
                float s1f[4] = {0.f, 0.f, 0.f, 0.f}, s2f[4] = {0.f, 0.f, 0.f, 0.f};
                float* orow = A.out + (size_t)(row0 + 32 * q) * NOUT + c0 + lane;
                int r = 0;
                if (keepbits == 0xffffffffu) {                                // no masked row in this 32-row block: the plain loop
                    for (; r + 4 <= nv; r += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            float o = fmaf(lds_f32(sc + 4 * ((r + u) * 36 + lane)), HALF ? H_WSCALE_INV : 1.f, bias_c);
                            o = fmaxf(o, o * LEAKY);
                            orow[(size_t)(r + u) * NOUT] = o;
                            s1f[u] += o; s2f[u] = fmaf(o, o, s2f[u]);
                        }
                    }
                }
                for (; r < nv; ++r) {
                    float o = fmaf(lds_f32(sc + 4 * (r * 36 + lane)), HALF ? H_WSCALE_INV : 1.f, bias_c);
                    o = fmaxf(o, o * LEAKY);
                    if (!((keepbits >> r) & 1u)) o = 0.f;                     // node mask: the row is stored as zero
                    orow[(size_t)r * NOUT] = o;
                    s1f[0] += o; s2f[0] = fmaf(o, o, s2f[0]);
                }
                if (A.part) {
                    sts_f32(red + 4 * ((q * 2 + 0) * NOUT + c0 + lane), (s1f[0] + s1f[1]) + (s1f[2] + s1f[3]));
                    sts_f32(red + 4 * ((q * 2 + 1) * NOUT + c0 + lane), (s2f[0] + s2f[1]) + (s2f[2] + s2f[3]));
                }
                __syncwarp();
            }
            // all tcgen05.ld of this accumulator are complete (wait::ld inside tmem_ld32): hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce + 8 * ab);
            if (A.part) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                double* p = A.part + (size_t)tile * 2 * NOUT;
                for (int i = etid; i < 2 * NOUT; i += 128) {
                    const int qq = i / NOUT, cc = i % NOUT;
                    p[i] = (((double)lds_f32(red + 4 * ((0 * 2 + qq) * NOUT + cc)) + (double)lds_f32(red + 4 * ((1 * 2 + qq) * NOUT + cc))) +
                            (double)lds_f32(red + 4 * ((2 * 2 + qq) * NOUT + cc))) + (double)lds_f32(red + 4 * ((3 * 2 + qq) * NOUT + cc));
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    if (A.dbg && blockIdx.x == 0 && lane == 0) { long long* d = A.dbg + warp * 4; d[0] = clock64() - t_start; d[1] = w0; d[2] = w1; }
    tc_fence_before();
    __syncthreads();
    if (timeout_flag && tid == 0) { *reinterpret_cast<volatile int*>(A.error_flag) = TGNN_DEVERR_PIPELINE; __threadfence_system(); }   // mapped host word
    if (warp == MMA_WARP) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// w: [N_out][K] (the checkpoint layout).  img: per 32-wide K slab, the hi tile then the lo tile, each the
// SWIZZLE_128B shared-memory image [N_out rows x 128 B] -- so a slab's B operand is one contiguous bulk copy.
__global__ void k_weight_image(const float* __restrict__ w, float* __restrict__ img, int n_out, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out * K) return;
    const int n = i / K, k = i - n * K, slab = k >> 5, kk = k & 31;
    const float x = w[i];
    const uint32_t h = tf32_rna(x);
    const uint32_t l = tf32_rna(x - __uint_as_float(h));
    const size_t base = (size_t)slab * 2 * n_out * 32;
    const int pos = n * 32 + ((((kk >> 2) ^ (n & 7)) << 2) | (kk & 3));
    img[base + pos] = __uint_as_float(h);
    img[base + (size_t)n_out * 32 + pos] = __uint_as_float(l);
}

// fp16 images: per 32-wide K slab ONE tile [N_out rows x 128 B], row n = {hi(w * 2^6)[32] | lo[32]} halves, SWIZZLE_128B.
__global__ void k_weight_image_h(const float* __restrict__ w, __half* __restrict__ img, int n_out, int K, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out * K) return;
    const int n = i / K, k = i - n * K, slab = k >> 5, kk = k & 31;
    const float x = w[i] * H_WSCALE;
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    if (!(fabsf(x) <= TG_H_LIMIT)) *flag = 1;
    __half* tile = img + (size_t)slab * n_out * 64;
    tile[n * 64 + ((((kk >> 3) ^ (n & 7)) << 3) | (kk & 7))] = hi;
    tile[n * 64 + ((((4 + (kk >> 3)) ^ (n & 7)) << 3) | (kk & 7))] = lo;
}

template <int NOUT, bool HALF>
void launch_one(const DenseTcArgs& a, int sm_count, cudaStream_t st) {
    constexpr size_t smem = DenseCfg<NOUT, HALF>::SMEM;
    static PerDeviceOnce once;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute((k_dense_tc<NOUT, HALF>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    });
    const int n_tiles = (a.n + BM - 1) / BM;
    static long long* dbg = nullptr;
    static const bool want_dbg = getenv("TGNN_DENSE_DBG") != nullptr;
    if (want_dbg && !dbg) TGNN_CUDA(cudaMalloc(&dbg, 16 * 4 * sizeof(long long)));
    DenseTcArgs b = a;
    b.dbg = dbg;
    k_dense_tc<NOUT, HALF><<<std::min(n_tiles, sm_count), NTHREADS, smem, st>>>(b);
    TGNN_CUDA(cudaGetLastError());
    if (want_dbg) {
        static int calls = 0;
        if (++calls == 8) {
            long long h[16 * 4];
            TGNN_CUDA(cudaStreamSynchronize(st));
            TGNN_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
            for (int w = 0; w < NTHREADS / 32; ++w)
                fprintf(stderr, "dense<%d> dbg warp %2d total %9lld wait0 %9lld wait1 %9lld\n", NOUT, w, h[4 * w], h[4 * w + 1], h[4 * w + 2]);
        }
    }
}

}  // namespace

void launch_weight_image(const float* w, float* img, int n_out, int K, cudaStream_t st) {
    k_weight_image<<<(n_out * K + 255) / 256, 256, 0, st>>>(w, img, n_out, K);
    TGNN_CUDA(cudaGetLastError());
}

int dense_tc_row_blocks(int n) { return (n + BM - 1) / BM; }

void launch_weight_image_h(const float* w, void* img, int n_out, int K, int* flag, cudaStream_t st) {
    k_weight_image_h<<<(n_out * K + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<__half*>(img), n_out, K, flag);
    TGNN_CUDA(cudaGetLastError());
}

template <bool HALF>
static void launch_dense_variant(const DenseTcArgs& a, int n_out, int sm_count, cudaStream_t st) {
    switch (n_out) {
        case 256: launch_one<256, HALF>(a, sm_count, st); break;
        case 128: launch_one<128, HALF>(a, sm_count, st); break;
        case 64: launch_one<64, HALF>(a, sm_count, st); break;
        case 32: launch_one<32, HALF>(a, sm_count, st); break;
        default: TGNN_CHECK(false, "dense stage: n_out must be 32, 64, 128 or 256");
    }
}

// w_img16 != null: the fp16 kernel runs first and the 3xTF32 kernel stands by behind it (range_flag, zeroed by the caller
// per forward); null: 3xTF32 only.  Returns the number of launches.
int launch_dense_tc(const DenseArgs& d, const float* w_img, const void* w_img16, int* range_flag, int* error_flag, int sm_count,
                    cudaStream_t st) {
    TGNN_CHECK(d.K % BK == 0, "dense stage: K must be a multiple of 32");
    DenseTcArgs a{};
    a.slabs = d.slabs; a.a = d.a; a.virtual_concat = d.virtual_concat; a.in_coef = d.in_coef;
    a.bias = d.bias; a.out = d.out; a.part = d.part; a.error_flag = error_flag; a.mask = d.mask;
    a.n = d.n; a.K = d.K; a.range_flag = range_flag;
    a.fin = d.fin;                           // the kernel that runs first finishes the input's BatchNorm and publishes in_coef ...
    if (w_img16) {
        a.w_img = reinterpret_cast<const float*>(w_img16); a.standby = 0;
        launch_dense_variant<true>(a, d.n_out, sm_count, st);
        a.fin = BnFin{};                     // ... the stand-by behind it reads the published coefficients
    }
    a.w_img = w_img; a.standby = w_img16 ? 1 : 0;
    launch_dense_variant<false>(a, d.n_out, sm_count, st);
    return w_img16 ? 2 : 1;
}

}  // namespace tgnn
