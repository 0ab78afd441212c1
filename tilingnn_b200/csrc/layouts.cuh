// Index maps of the pre-arranged weight tables (shared by the kernels that read them and tables.cu that writes them).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tgnn {

// ---- 3xTF32 fragment tables of mma.sync m16n8k8 (k_conv_adj, k_gin; kernels.cu) -------------------------------
//  KMAP_GATHER : A comes from two float4 loads per row (cols 4t.., 16+4t..): k-step ks, slot kk=tt+4e  <-> col 16(ks>>1)+4tt+2(ks&1)+e
//  KMAP_NATURAL: A comes from shared memory, col = 8ks + kk
//  KMAP_CHAIN  : A is the previous layer's C fragment: k-step ks = previous n-tile, slot kk=tt+4e <-> unit 8ks + 2tt + e
//  NMAP_NATURAL: unit = 8nt + g          NMAP_CONTIG8: channel = 8(g>>1) + 2nt + (g&1)   (N = 32 only)
enum { KMAP_GATHER = 0, KMAP_NATURAL = 1, KMAP_CHAIN = 2, NMAP_NATURAL = 0, NMAP_CONTIG8 = 1 };

__device__ __forceinline__ uint32_t tf32_rna_bits(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__host__ __device__ __forceinline__ void frag_coords(int k, int n, int kmap, int nmap, int& ks, int& tt, int& e, int& nt, int& g) {
    if (kmap == KMAP_GATHER) { int r = k & 15; ks = 2 * (k >> 4) + ((r >> 1) & 1); tt = r >> 2; e = r & 1; }
    else if (kmap == KMAP_NATURAL) { ks = k >> 3; int kk = k & 7; tt = kk & 3; e = kk >> 2; }
    else { ks = k >> 3; int r = k & 7; tt = r >> 1; e = r & 1; }
    if (nmap == NMAP_NATURAL) { nt = n >> 3; g = n & 7; }
    else { int r = n & 7; nt = r >> 1; g = 2 * (n >> 3) + (r & 1); }
}
// float index of element (k, n), hi (hl=0) or lo (hl=1) part, in a frag table of an [K x N] matrix:
//   float4 index ((ks*2 + hl) * (N/16) + j) * 32 + lane ; float4 = {b0,b1 of n-tile 2j, b0,b1 of n-tile 2j+1}
__host__ __device__ __forceinline__ size_t frag_index(int k, int n, int N, int kmap, int nmap, int hl) {
    int ks, tt, e, nt, g;
    frag_coords(k, n, kmap, nmap, ks, tt, e, nt, g);
    const int lane = g * 4 + tt, j = nt >> 1, comp = 2 * (nt & 1) + e;
    return ((((size_t)ks * 2 + hl) * (N / 16) + j) * 32 + lane) * 4 + comp;
}
__device__ __forceinline__ void frag_store(float* tab, int k, int n, int N, int kmap, int nmap, double w) {
    const uint32_t hi = tf32_rna_bits((float)w);
    const uint32_t lo = tf32_rna_bits((float)(w - (double)__uint_as_float(hi)));
    tab[frag_index(k, n, N, kmap, nmap, 0)] = __uint_as_float(hi);
    tab[frag_index(k, n, N, kmap, nmap, 1)] = __uint_as_float(lo);
}

// ---- tcgen05 S kernel (conv_s.cu): float index of element (row n, column k) inside the SWIZZLE_128B shared-memory
// IMAGE of a [32 x 32] tf32 tile; the tables are stored pre-swizzled so one bulk copy drops a ready operand tile ----
__host__ __device__ __forceinline__ int tile_pos(int n, int k) { return n * 32 + ((((k >> 2) ^ (n & 7)) << 2) | (k & 3)); }

// ---- fp16 hi|lo fragment table of mma.sync m16n8k16 (k_conv_h; conv_h.cu): 2048 halves = 1024 words --------------
//   k -> k16 step ks = k>>4, thread-in-group tt = (k&15)>>2, register (k&3)>>1, half k&1   (matches the row layout of xh)
//   n -> n-tile nt = (n&7)>>1, group g = 2(n>>3) + (n&1)   (a lane ends up with 8 contiguous output channels 8t..8t+7)
//   uint4 index ((ks*2 + hl)*2 + j)*32 + lane, lane = 4g + tt, j = nt>>1; component 2(nt&1) + register
__host__ __device__ __forceinline__ int hfrag_half_index(int k, int n, int hl) {
    const int ks = k >> 4, r = k & 15, tt = r >> 2, reg = (r & 3) >> 1, e = r & 1;
    const int rn = n & 7, nt = rn >> 1, g = 2 * (n >> 3) + (rn & 1);
    const int lane = g * 4 + tt, j = nt >> 1, comp = 2 * (nt & 1) + reg;
    return (((((ks * 2 + hl) * 2 + j) * 32 + lane) * 4 + comp) * 2) + e;
}
__device__ __forceinline__ void hfrag_store(__half* tab, int k, int n, float w, int* flag, float limit) {
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
    tab[hfrag_half_index(k, n, 0)] = hi;
    tab[hfrag_half_index(k, n, 1)] = lo;
    if (flag && !(fabsf(w) <= limit)) *flag = 1;
}

// ---- fp16 hi|lo A-operand fragment table of W^T for the transposed-roles kernel (k_conv_x; conv_h.cu) ----------------
//   MMA row r of channel half i  <-> output channel 16 i + 2 (r & 7) + (r >> 3)      (rows g / g+8 = adjacent channels)
//   MMA k index kappa of k16 step ks <-> input channel 16 ks + 4 ((kappa & 7) >> 1) + 2 (kappa >> 3) + (kappa & 1)
//     (= the order of a split row's storage: a lane's B registers are consecutive words of its 256-bit row load)
//   uint4 index ((i*2 + ks)*2 + hl)*32 + lane, lane = 4 (r & 7) + ((kappa & 7) >> 1); register (r >> 3) + 2 (kappa >> 3); half kappa & 1
__host__ __device__ __forceinline__ int xfrag_half_index(int kin, int n, int hl) {
    const int ks = kin >> 4, rem = kin & 15, tp = rem >> 2, within = rem & 3;
    const int kappa = within < 2 ? 2 * tp + within : 8 + 2 * tp + (within - 2);
    const int i = n >> 4, nr = n & 15, r = (nr & 1) ? 8 + (nr >> 1) : (nr >> 1);
    const int lane = 4 * (r & 7) + ((kappa & 7) >> 1), reg = (r >> 3) + 2 * (kappa >> 3), e = kappa & 1;
    return ((((i * 2 + ks) * 2 + hl) * 32 + lane) * 4 + reg) * 2 + e;
}
__device__ __forceinline__ void xfrag_store(__half* tab, int kin, int n, float w, int* flag, float limit) {
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
    tab[xfrag_half_index(kin, n, 0)] = hi;
    tab[xfrag_half_index(kin, n, 1)] = lo;
    if (flag && !(fabsf(w) <= limit)) *flag = 1;
}

// ---- tcgen05 edge-block kernel (conv_t.cu): the [64 rows n][64 halves k] K-major SWIZZLE_128B IMAGE of a type's weights.
// k runs in the storage order of a split activation row (hsplit.cuh): 16-byte piece p = xh_pos(c >> 2) holds
// {hi(4q..4q+3), lo(4q..4q+3)} of channels 4q + i.  Rows  0..31 ("main",  out channel n): hi-slot = Whi, lo-slot = 0;
// rows 32..63 ("small", out channel n - 32): hi-slot = Wlo (scaled 2^11), lo-slot = Whi  ->  D = A . B^T gives
//   D[:, n] = hi . Whi      D[:, 32 + n] = hi . Wlo + lo . Whi      message = D[:, n] + 2^-11 D[:, 32 + n].
__host__ __device__ __forceinline__ int timg_half_index(int row, int k) { return row * 64 + ((((k >> 3) ^ (row & 7)) << 3) | (k & 7)); }
__device__ __forceinline__ void timg_store(__half* img, int kin, int n, float w, int* flag, float limit) {
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
    const int q = kin >> 2, i = kin & 3, p = 2 * (q & 3) + (q >> 2);          // p = xh_pos(q)
    const int k_hi = 8 * p + i, k_lo = 8 * p + 4 + i;
    img[timg_half_index(n, k_hi)] = hi;
    img[timg_half_index(n, k_lo)] = __float2half_rn(0.f);
    img[timg_half_index(32 + n, k_hi)] = lo;
    img[timg_half_index(32 + n, k_lo)] = hi;
    if (flag && !(fabsf(w) <= limit)) *flag = 1;
}

// ---- fp16 hi|lo fragment table of a k-major [K][N] matrix in NATURAL k order (k = 16 ks + 8 reg + 2 tt + e), for
// the chained GIN layers whose A fragments are the previous layer's C fragments (k_gin<true>; kernels.cu) ----------
//   uint4 index ((ks*2 + hl) * (N/16) + j)*32 + lane ; component 2(nt&1) + reg ; half e
__host__ __device__ __forceinline__ size_t hfrag_nat_half_index(int k, int n, int N, int nmap, int hl) {
    const int ks = k >> 4, r = k & 15, reg = r >> 3, tt = (r & 7) >> 1, e = r & 1;
    int nt, g;
    if (nmap == NMAP_NATURAL) { nt = n >> 3; g = n & 7; }
    else { const int r8 = n & 7; nt = r8 >> 1; g = 2 * (n >> 3) + (r8 & 1); }
    const int lane = g * 4 + tt, j = nt >> 1, comp = 2 * (nt & 1) + reg;
    return ((((((size_t)ks * 2 + hl) * (N / 16) + j) * 32 + lane) * 4 + comp) * 2) + e;
}

}  // namespace tgnn
