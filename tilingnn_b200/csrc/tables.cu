// Per-type edge-weight tables of the adjacency branch, all layers in ONE launch.
//
// The reference runs a 3-layer sigmoid MLP  D_e -> 32 -> 64 -> 1024  on every adjacency edge's feature row
// (graph_networks/layers/edge_conv.py:17-18, util.py:10-17) and reshapes the result to W_e[32 in][32 out]
// (PyG NNConv: weight.view(-1, in, out)).  Equal rows give equal weights, so the MLP is evaluated once per distinct
// row ("edge type") and layer -- in fp64, rounded once to fp32 -- and stored in the layouts the three adjacency
// kernels read (layouts.cuh): 3xTF32 fragments (k_conv_adj), the pre-swizzled tcgen05 operand image (k_conv_s) and
// fp16 hi|lo fragments (k_conv_h), the pre-swizzled [64 x 64] fp16 image of the tcgen05 edge-block kernel (k_conv_t) and plain
// fp32 (its wide-range stand-in).  Entry K of every layer is nnConv.root.  grid = (K + 1, L), block = 256.
#include "layouts.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {

__global__ void __launch_bounds__(256)
k_edge_tables(const float* __restrict__ rows, int K, int d_e, const TableLayer* __restrict__ layers,
              float* __restrict__ tabF, float* __restrict__ tabS, uint32_t* __restrict__ tabH, uint32_t* __restrict__ tabT,
              float* __restrict__ tab32, int* __restrict__ wflags, uint32_t* __restrict__ tabX) {
    __shared__ double h1[32], h2[64];
    const int t = blockIdx.x, layer = blockIdx.y, tid = threadIdx.x;
    const TableLayer L = layers[layer];
    const size_t slot = (size_t)layer * (K + 1) + t;
    float* outF = tabF ? tabF + slot * TG_FRAG32 : nullptr;
    float* outS = tabS ? tabS + slot * 2048 : nullptr;
    __half* outH = tabH ? reinterpret_cast<__half*>(tabH + slot * TG_HFRAG32) : nullptr;
    __half* outT = tabT ? reinterpret_cast<__half*>(tabT + slot * TG_TIMG32) : nullptr;
    float* out32 = tab32 ? tab32 + slot * (F * F) : nullptr;
    __half* outX = tabX ? reinterpret_cast<__half*>(tabX + slot * TG_HFRAG32) : nullptr;
    const bool is_root = t == K;
    if (!is_root) {
        const float* e = rows + (size_t)t * d_e;
        if (tid < 32) {
            double s = (double)L.c1[tid];
            for (int k = 0; k < d_e; ++k) s += (double)L.a1[tid * d_e + k] * (double)e[k];
            h1[tid] = 1.0 / (1.0 + exp(-s));
        }
        __syncthreads();
        if (tid < 64) {
            double s = (double)L.c2[tid];
            for (int k = 0; k < 32; ++k) s += (double)L.a2[tid * 32 + k] * h1[k];
            h2[tid] = 1.0 / (1.0 + exp(-s));
        }
        __syncthreads();
    }
    for (int o = tid; o < F * F; o += 256) {
        double w;
        if (is_root) w = (double)L.root[o];                                   // nnConv.root is [in][out] already
        else {
            double s = (double)L.c3[o];
            for (int k = 0; k < 64; ++k) s += (double)L.a3[(size_t)o * 64 + k] * h2[k];
            w = 1.0 / (1.0 + exp(-s));
        }
        const int kin = o >> 5, n = o & 31;                                   // o = k_in * 32 + k_out
        if (outF) frag_store(outF, kin, n, 32, KMAP_GATHER, NMAP_CONTIG8, w);
        if (outS) {
            const uint32_t hi = tf32_rna_bits((float)w);
            const uint32_t lo = tf32_rna_bits((float)(w - (double)__uint_as_float(hi)));
            const int pos = tile_pos(n, kin);
            outS[pos] = __uint_as_float(hi);
            outS[1024 + pos] = __uint_as_float(lo);
        }
        // a sigmoid is always inside the fp16 range; a root weight may not be (then k_conv_adj takes the layer)
        if (outH) hfrag_store(outH, kin, n, (float)w, is_root ? wflags + layer : nullptr, TG_H_LIMIT);
        if (outT) timg_store(outT, kin, n, (float)w, is_root ? wflags + layer : nullptr, TG_H_LIMIT);
        if (out32) out32[o] = (float)w;
        if (outX) xfrag_store(outX, kin, n, (float)w, is_root ? wflags + layer : nullptr, TG_H_LIMIT);
    }
}

}  // namespace

void launch_edge_tables(const float* type_rows, int n_types, int d_e, int n_layers, const TableLayer* layers_dev,
                        float* tabF, float* tabS, uint32_t* tabH, uint32_t* tabT, float* tab32, int* wflags, cudaStream_t st,
                        uint32_t* tabX) {
    k_edge_tables<<<dim3(n_types + 1, n_layers), 256, 0, st>>>(type_rows, n_types, d_e, layers_dev, tabF, tabS, tabH, tabT, tab32, wflags, tabX);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
