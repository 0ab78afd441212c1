// Dense stages of the final MLP (TilinGNN.py:45-46,74-76 of the reference) on the 5th-generation
// tensor cores:  out = LeakyReLU( BN_in(A) @ W^T + b )  and per-column BatchNorm partial sums.
//
// Persistent CTAs (one per SM) walk tiles of 128 rows x all N_out columns.  The fp32 accumulator lives in
// TMEM (128 lanes x N_out columns), double buffered so the epilogue of one tile overlaps the next main loop.  K is walked in slabs of 32 (= one 128-byte SWIZZLE_128B atom row):
//   producers (warps 0-7): A slab  global -> registers -> lazy BatchNorm -> hi/lo TF32 split -> swizzled smem
//                          W slab  (pre-split hi / lo, [N_out][K] = the checkpoint's own layout, K-major)
//                                  cp.async -> swizzled smem
//   MMA issuer (warp 8, one lane): 12 x tcgen05.mma.kind::tf32 per slab (4 K-steps of 8 x {lo*hi, hi*lo, hi*hi}),
//                          tcgen05.commit -> mbarrier frees the smem stage / publishes the accumulator
//   epilogue (warps 9-12): tcgen05.ld 32 columns at a time -> bias, LeakyReLU -> global, column sums in fp64.
// 3xTF32 keeps fp32-level accuracy (single-pass TF32 would break the 1e-4 parity bar).
// Every mbarrier wait is bounded: on a timeout the kernel raises a device-side error flag and exits instead of
// hanging the GPU.
#include <algorithm>

#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {

constexpr int BM = 128;           // rows per tile (UMMA M)
constexpr int BK = 32;            // K per slab (128 bytes of tf32)
constexpr int A_TILE_BYTES = BM * BK * 4;            // 16 KB
constexpr int N_PROD_WARPS = 8, MMA_WARP = 8, EPI_WARP0 = 9;
constexpr int NTHREADS = 13 * 32;                    // 8 producer warps, 1 MMA warp, 4 epilogue warps
constexpr int SCRATCH_FLOATS = 32 * 33;              // per epilogue warp: one 32x32 block, padded

using namespace tc;

struct DenseTcArgs {
    const float* const* slabs;  // virtual concat: K/32 slab pointers [n_rows][32]
    const float* a;             // else plain [n][K]
    int virtual_concat;
    const float* in_coef;       // [4][K] lazy BatchNorm of the input, or nullptr
    const float* w_hi;          // [N_out][K] tf32-rounded weights
    const float* w_lo;          // [N_out][K] tf32-rounded residuals
    const float* bias;          // [N_out]
    float* out;                 // [n][N_out]
    double* part;               // [gridDim.x][2][N_out] or nullptr
    int* error_flag;
    int n, K;
};

template <int NOUT> struct DenseCfg {
    static constexpr int B_TILE_BYTES = NOUT * BK * 4;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = NOUT >= 256 ? 2 : (NOUT >= 128 ? 3 : 4);
    static constexpr int EPI_BYTES = 4 * SCRATCH_FLOATS * 4 + 4 * 2 * NOUT * 8 + NOUT * 4;   // scratch, red (fp64), bias
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024;
};

// Persistent, warp-specialised:  producers (warps 0-7) -> smem ring -> MMA warp -> TMEM (double buffered)
// -> epilogue warps (9-12).  The epilogue of tile i overlaps the main loop of tile i+1.
template <int NOUT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_dense_tc(DenseTcArgs A) {
    using Cfg = DenseCfg<NOUT>;
    constexpr int STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES, B_TILE_BYTES = Cfg::B_TILE_BYTES;
    constexpr uint32_t IDESC = umma_idesc_tf32(NOUT);
    constexpr int TMEM_COLS = 2 * NOUT < 32 ? 32 : 2 * NOUT;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int timeout_flag;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t epi = smem_base + STAGES * STAGE_BYTES;
    const uint32_t scratch_all = epi;                                                // [4][32*33] float
    const uint32_t red = epi + 4 * SCRATCH_FLOATS * 4;                               // [4][2][NOUT] double
    const uint32_t bias_s = epi + 4 * SCRATCH_FLOATS * 4 + 4 * 2 * NOUT * 8;         // [NOUT] float
    const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[STAGES]);
    const uint32_t bar_accf = smem_u32(&bars[2 * STAGES]), bar_acce = smem_u32(&bars[2 * STAGES + 2]);
    const int n_slabs = A.K / BK;
    const int n_tiles = (A.n + BM - 1) / BM;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, N_PROD_WARPS); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_accf + 8 * b, 1); mbar_init(bar_acce + 8 * b, 4); }
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

    if (warp < N_PROD_WARPS) {
        // ===================== producers: 256 threads, item (row = tid/8 + 32 j, chunk = tid%8) =====================
        const int c = tid & 7, rbase = tid >> 3;
        auto load_a = [&](int tile, int s, float4 (&v)[4]) {
            const int row0 = tile * BM;
            const float* abase; size_t lda; int koff;
            if (A.virtual_concat) { abase = A.slabs[s]; lda = F; koff = 0; }
            else { abase = A.a; lda = (size_t)A.K; koff = s * BK; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = row0 + rbase + 32 * j;
                v[j] = r < A.n ? __ldg(reinterpret_cast<const float4*>(abase + (size_t)r * lda + koff) + c)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        float4 pre[4];
        int g = 0;
        bool ok = true;
        if ((int)blockIdx.x < n_tiles) load_a(blockIdx.x, 0, pre);
        for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
            const int row0 = tile * BM;
            for (int s = 0; s < n_slabs; ++s, ++g) {
                const int st = g % STAGES;
                if (!mbar_wait(bar_empty + 8 * st, ((g / STAGES) & 1) ^ 1)) { timeout_flag = 1; ok = false; break; }
                const uint32_t sa_hi = smem_base + st * STAGE_BYTES, sa_lo = sa_hi + A_TILE_BYTES;
                const uint32_t sb_hi = smem_base + st * STAGE_BYTES + 2 * A_TILE_BYTES;
                const uint32_t sb_lo = sb_hi + B_TILE_BYTES;
                const int k0 = s * BK;
                for (int i = tid; i < NOUT * 8; i += N_PROD_WARPS * 32) {
                    const int r = i >> 3, cc = i & 7;
                    const uint32_t off = sw128_off(r, cc);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb_hi + off), "l"(A.w_hi + (size_t)r * A.K + k0 + 4 * cc) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb_lo + off), "l"(A.w_lo + (size_t)r * A.K + k0 + 4 * cc) : "memory");
                }
                float4 cur[4] = {pre[0], pre[1], pre[2], pre[3]};
                // prefetch the next slab's rows (possibly the next tile's first slab) while this one is transformed
                {
                    int nt = tile, ns = s + 1;
                    if (ns == n_slabs) { ns = 0; nt = tile + gridDim.x; }
                    if (nt < n_tiles) load_a(nt, ns, pre);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = rbase + 32 * j;
                    float4 v = cur[j];
                    if (A.in_coef) {
                        const float* cf = A.in_coef + k0 + 4 * c;
                        const int C = A.K;
                        v.x = fmaf((v.x - cf[0]) - cf[C + 0], cf[2 * C + 0], cf[3 * C + 0]);
                        v.y = fmaf((v.y - cf[1]) - cf[C + 1], cf[2 * C + 1], cf[3 * C + 1]);
                        v.z = fmaf((v.z - cf[2]) - cf[C + 2], cf[2 * C + 2], cf[3 * C + 2]);
                        v.w = fmaf((v.w - cf[3]) - cf[C + 3], cf[2 * C + 3], cf[3 * C + 3]);
                        if (row0 + r >= A.n) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    uint4 hi, lo;
                    hi.x = tf32_rna(v.x); lo.x = tf32_rna(v.x - __uint_as_float(hi.x));
                    hi.y = tf32_rna(v.y); lo.y = tf32_rna(v.y - __uint_as_float(hi.y));
                    hi.z = tf32_rna(v.z); lo.z = tf32_rna(v.z - __uint_as_float(hi.z));
                    hi.w = tf32_rna(v.w); lo.w = tf32_rna(v.w - __uint_as_float(hi.w));
                    const uint32_t off = sw128_off(r, c);
                    sts128(sa_hi + off, hi);
                    sts128(sa_lo + off, lo);
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                fence_proxy_async();                   // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * st);
            }
        }
    } else if (warp == MMA_WARP) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            int g = 0, it = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
                const int ab = it & 1;
                if (!mbar_wait(bar_acce + 8 * ab, ((it >> 1) & 1) ^ 1)) { timeout_flag = 1; break; }
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ab * NOUT);
                for (int s = 0; s < n_slabs; ++s, ++g) {
                    const int st = g % STAGES;
                    if (!mbar_wait(bar_full + 8 * st, (g / STAGES) & 1)) { timeout_flag = 1; ok = false; break; }
                    tc_fence_after();
                    const uint32_t a_hi = smem_base + st * STAGE_BYTES, a_lo = a_hi + A_TILE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_TILE_BYTES, b_lo = b_hi + B_TILE_BYTES;
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
        const int q = warp & 3, etid = (warp - EPI_WARP0) * 32 + lane;
        const uint32_t sc = scratch_all + (uint32_t)q * SCRATCH_FLOATS * 4;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            if (!mbar_wait(bar_accf + 8 * ab, (it >> 1) & 1)) { timeout_flag = 1; break; }
            tc_fence_after();
            const int row0 = tile * BM, row = row0 + 32 * q + lane;
            const bool live = row < A.n;
            int nv = A.n - (row0 + 32 * q);
            nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
#pragma unroll 1
            for (int c0 = 0; c0 < NOUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(ab * NOUT + c0), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(leaky(__uint_as_float(v[j]) + lds_f32(bias_s + 4 * (c0 + j))));
                if (live) {
                    float4* dst = reinterpret_cast<float4*>(A.out + (size_t)row * NOUT + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
                if (A.part) {
                    // column sums over this warp's rows through a padded 32x32 scratch block (lane = column)
#pragma unroll
                    for (int j = 0; j < 32; ++j) sts_f32(sc + 4 * (lane * 33 + j), __uint_as_float(v[j]));
                    __syncwarp();
                    double s1 = 0.0, s2 = 0.0;
                    for (int r = 0; r < nv; ++r) {
                        const double o = (double)lds_f32(sc + 4 * (r * 33 + lane));
                        s1 += o; s2 += o * o;
                    }
                    sts_f64(red + 8 * ((q * 2 + 0) * NOUT + c0 + lane), s1);
                    sts_f64(red + 8 * ((q * 2 + 1) * NOUT + c0 + lane), s2);
                    __syncwarp();
                }
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
                    p[i] = ((lds_f64(red + 8 * ((0 * 2 + qq) * NOUT + cc)) + lds_f64(red + 8 * ((1 * 2 + qq) * NOUT + cc))) +
                            lds_f64(red + 8 * ((2 * 2 + qq) * NOUT + cc))) + lds_f64(red + 8 * ((3 * 2 + qq) * NOUT + cc));
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (timeout_flag && tid == 0) atomicExch(A.error_flag, 1);
    if (warp == MMA_WARP) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

__global__ void k_split_tf32(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = w[i];
    const uint32_t h = tf32_rna(x);
    hi[i] = __uint_as_float(h);
    lo[i] = __uint_as_float(tf32_rna(x - __uint_as_float(h)));
}

template <int NOUT>
void launch_one(const DenseTcArgs& a, int sm_count, cudaStream_t st) {
    constexpr size_t smem = DenseCfg<NOUT>::SMEM;
    static bool attr = false;
    if (!attr) {
        TGNN_CUDA(cudaFuncSetAttribute(k_dense_tc<NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int n_tiles = (a.n + BM - 1) / BM;
    k_dense_tc<NOUT><<<std::min(n_tiles, sm_count), NTHREADS, smem, st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace

void launch_split_tf32(const float* w, float* hi, float* lo, int n, cudaStream_t st) {
    k_split_tf32<<<(n + 255) / 256, 256, 0, st>>>(w, hi, lo, n);
    TGNN_CUDA(cudaGetLastError());
}

int dense_tc_row_blocks(int n) { return (n + BM - 1) / BM; }

void launch_dense_tc(const DenseArgs& d, const float* w_hi, const float* w_lo, int* error_flag, int sm_count, cudaStream_t st) {
    TGNN_CHECK(d.K % BK == 0, "dense stage: K must be a multiple of 32");
    DenseTcArgs a{};
    a.slabs = d.slabs; a.a = d.a; a.virtual_concat = d.virtual_concat; a.in_coef = d.in_coef;
    a.w_hi = w_hi; a.w_lo = w_lo; a.bias = d.bias; a.out = d.out; a.part = d.part; a.error_flag = error_flag;
    a.n = d.n; a.K = d.K;
    switch (d.n_out) {
        case 256: launch_one<256>(a, sm_count, st); break;
        case 128: launch_one<128>(a, sm_count, st); break;
        case 64: launch_one<64>(a, sm_count, st); break;
        case 32: launch_one<32>(a, sm_count, st); break;
        default: TGNN_CHECK(false, "dense stage: n_out must be 32, 64, 128 or 256");
    }
}

}  // namespace tgnn
