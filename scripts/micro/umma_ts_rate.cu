// Microbenchmark: cycles per tcgen05.mma with the A operand in TENSOR MEMORY (M=128), issued back-to-back by one thread:
// kind::tf32 (K=8) / kind::f16 (K=16), N = 32 / 64 / 128, accumulating into one D tile or rotating over ND tiles,
// and the shared-memory-A form beside it.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../tilingnn_b200/csrc/tc_common.cuh"
using namespace tgnn::tc;
__host__ __device__ constexpr uint32_t idesc_f16(int n) {   // D = F32, A = B = F16, K-major, M = 128
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, bool f16) {
    if (f16) asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(db), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(db), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, bool f16) {
    if (f16) asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
template <int N, int ND, bool F16, bool TS, int STORM>
__global__ void k(long long* out, int iters) {
    __shared__ volatile int stop;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = smem_u32(smem);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_s;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (STORM && threadIdx.x >= 128) {
        // 4 warps (one per TMEM lane quarter) store 32 columns per instruction into columns 384..447, as the gather warps do
        const uint32_t ta = tm + ((uint32_t)(32 * ((threadIdx.x >> 5) & 3)) << 16) + 384u;
        uint32_t v = threadIdx.x;
        long long n = 0;
        while (!stop) {
            for (int r = 0; r < STORM; ++r) {
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta + 32u * (r & 1)), "r"(v) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            ++n;
            if (STORM < 100) __nanosleep(STORM == 1 ? 400 : 100);
        }
        if ((threadIdx.x & 31) == 0) out[2 + ((threadIdx.x >> 5) & 3)] = n;
    }
    if (threadIdx.x == 0) {
        const uint64_t da = umma_desc_sw128(sb), db = umma_desc_sw128(sb + 16384);
        const uint32_t idesc = F16 ? idesc_f16(N) : umma_idesc_tf32(N);
        const uint32_t a0 = tm + 448;                       // A operand columns (content irrelevant for timing)
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t d = tm + (uint32_t)(((4 * i + j) % ND) * N);
                if (TS) mma_ts(d, a0 + 8 * j, db + 2 * j, idesc, F16); else mma_ss(d, da + 2 * j, db + 2 * j, idesc, F16);
            }
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
        stop = 1;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512)); }
}
template <int N, int ND, bool F16, bool TS, int STORM = 0> void run() {
    long long* d; cudaMalloc(&d, 64); long long h[8];
    cudaFuncSetAttribute(k<N, ND, F16, TS, STORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    const int iters = 2000;
    k<N, ND, F16, TS, STORM><<<1, STORM ? 256 : 128, 60000>>>(d, iters); k<N, ND, F16, TS, STORM><<<1, STORM ? 256 : 128, 60000>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    if (STORM) printf("[tcgen05.st storm %d: %.1f x32-stores per 1000 cycles per warp] ", STORM, 1000.0 * (double)h[2] * (STORM < 100 ? STORM : STORM) / (double)h[1]);
    printf("%s %s N=%3d D tiles=%d  issue %.1f cyc/mma   issue+drain %.1f cyc/mma   (%s)\n", F16 ? "f16 " : "tf32", TS ? "A=TMEM" : "A=smem", N, ND,
           (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run<32, 1, false, false>(); run<32, 4, false, false>();
    run<32, 1, false, true>(); run<32, 2, false, true>(); run<32, 4, false, true>(); run<64, 1, false, true>(); run<64, 4, false, true>(); run<128, 1, false, true>();
    run<32, 1, true, false>(); run<32, 1, true, true>(); run<32, 4, true, true>(); run<64, 1, true, true>(); run<64, 4, true, true>(); run<128, 1, true, true>();
    run<32, 1, false, true, 1>(); run<32, 1, false, true, 2>(); run<32, 1, false, true, 100>();
    run<64, 1, true, true, 1>(); run<64, 1, true, true, 2>(); run<64, 1, true, true, 100>();
    return 0;
}
