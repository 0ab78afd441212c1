// Microbenchmark: cycles per tcgen05.mma.kind::tf32 (M=128, K=8) issued back-to-back by one thread, vs N.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../tilingnn_b200/csrc/tc_common.cuh"
using namespace tgnn::tc;
template <int N>
__global__ void k(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sb = smem_u32(smem);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_s;
    if (threadIdx.x == 0) {
        const uint64_t da = umma_desc_sw128(sb), db = umma_desc_sw128(sb + 16384);
        const uint32_t idesc = umma_idesc_tf32(N);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            umma_tf32_acc(tm, da, db, idesc); umma_tf32_acc(tm, da + 2, db + 2, idesc);
            umma_tf32_acc(tm, da + 4, db + 4, idesc); umma_tf32_acc(tm, da + 6, db + 6, idesc);
        }
        long long t1 = clock64();
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256)); }
}
template <int N> void run() {
    long long* d; cudaMalloc(&d, 16); long long h[2];
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    const int iters = 2000;
    k<N><<<1, 128, 60000>>>(d, iters); k<N><<<1, 128, 60000>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d  issue %.1f cyc/mma   issue+drain %.1f cyc/mma   (%s)\n", N, (double)h[0] / (4.0 * iters), (double)h[1] / (4.0 * iters), cudaGetErrorString(e));
    cudaFree(d);
}
int main() { run<32>(); run<64>(); run<128>(); run<256>(); return 0; }
