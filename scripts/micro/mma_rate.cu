// Microbenchmark: issue rate of legacy mma.sync shapes on sm_100a (cycles per instruction per SM sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int MODE>
__global__ void k(float* out, int iters) {
    float c[8][4] = {};
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f801000u, 0x3f802000u, 0x3f803000u}, b0 = 0x3f800000u, b1 = 0x3f900000u;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (MODE == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (MODE == 3)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(b0));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0);
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int MODE> void run(const char* name, int warps_per_sm) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    int iters = 2000;
    k<MODE><<<148, 32 * warps_per_sm>>>(out, iters);
    k<MODE><<<148, 32 * warps_per_sm>>>(out, iters);
    cudaDeviceSynchronize();
    float cyc; cudaMemcpy(&cyc, out, 4, cudaMemcpyDeviceToHost);
    double per_smsp = (double)cyc / (iters * 8.0 * warps_per_sm / 4.0);
    printf("%-28s warps/SM=%2d  cycles per mma per SMSP = %6.2f\n", name, warps_per_sm, per_smsp);
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>("m16n8k8  tf32", w); run<4>("m16n8k4  tf32", w); run<1>("m16n8k16 bf16", w); run<2>("m16n8k16 f16", w); run<3>("m16n8k8  bf16", w);
    }
    return 0;
}
