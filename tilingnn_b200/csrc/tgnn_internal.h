// Internal declarations shared by the translation units of libtgnn.so.
// Not part of the public ABI (that is include/tgnn.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "tgnn.h"

namespace tgnn {

constexpr int F = 32;            // network_width (inputs/config.py:38 of the reference)
constexpr int WN_SMALL = 64;     // destination rows owned by one warp tile of the typed adjacency format ...
constexpr int WN_BIG = 128;      // ... and for large graphs: longer same-type runs (fewer weight-table reloads, less padding)
constexpr int CH = 16;           // edge slots per chunk (all of one edge type)
constexpr int GRP = 8;           // slots per group; destinations are distinct inside a group
constexpr int MAX_TYPES = 1 << 22; // distinct adjacency feature rows; each costs a 12 KB weight table per resident layer
constexpr float LEAKY = 0.01f;
constexpr double BN_EPS = 1e-5;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define TGNN_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            throw ::tgnn::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +   \
                                __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
    } while (0)

#define TGNN_CHECK(cond, msg)                                                                 \
    do {                                                                                      \
        if (!(cond)) throw ::tgnn::Error(std::string(msg));                                   \
    } while (0)

// cudaFuncSetAttribute is per DEVICE: a launcher keeps one of these (static) and prepares each device once,
// also when handles on several devices / threads share the process.
struct PerDeviceOnce {
    std::mutex m;
    uint64_t done = 0;                 // bit per device ordinal
    template <class Fn> void run(Fn&& fn) {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(m);
        if (dev < 64 && ((done >> dev) & 1ull)) return;
        fn();
        if (dev < 64) done |= 1ull << dev;
    }
};

// Owning device buffer (grow-only reuse keeps cudaMalloc out of repeated set_graph calls).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        TGNN_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// BatchNorm coefficients as consumers apply them:  y = ((x - mu_hi) - mu_lo) * scale + beta
// layout [4][C] floats: mu_hi, mu_lo, scale, beta.
struct BnRef {
    const float* coef = nullptr;   // device, [4][C]
};

// ----- graph structures (device) -------------------------------------------------------------
struct Graph {
    int64_t n_own = 0, n_rows = 0, n_global = 0;
    int64_t e_adj = 0, e_col = 0;
    int n_types = 0;
    int wn = WN_SMALL;        // rows per warp tile (WN_SMALL or WN_BIG)
    int n_tiles = 0;          // ceil(n_own / wn)
    int n_chunks = 0;
    // typed adjacency tiles
    DevBuf cptr;              // int32 [n_tiles + 1]   chunk range per warp tile
    DevBuf ctype;             // int32 [n_chunks]
    DevBuf csrc;              // int32 [n_chunks * CH] source row, -1 = empty slot
    DevBuf cdst;              // uint8 [n_chunks * CH] local destination row in the tile
    DevBuf inv_deg;           // float [n_own]         1 / max(1, in-degree)
    DevBuf type_rows;         // float [n_types * d_e] representative feature row per type
    // "S" format for the tcgen05 adjacency kernel: destinations in tiles of 128 rows; a tile's in-edges sorted by
    // (type, destination row); one PASS per distinct type present in the tile
    bool has_s = false;       // usable by k_conv_s (passes short enough for its row ring)
    bool s_built = false;     // the arrays below exist (k_conv_z needs them too)
    int s_tiles = 0, s_passes = 0, s_max_pass = 0;   // s_max_pass: most edges in one pass
    DevBuf s_pptr;            // int32  [s_tiles + 1]      pass range per tile
    DevBuf s_ptype;           // int32  [s_passes]         edge type of the pass
    DevBuf s_pbase;           // int32  [s_passes + 1]     first edge of the pass in s_src
    DevBuf s_off;             // uint16 [s_passes * 136]   per pass: edge offset of row r (r = 0..128) inside the pass
    DevBuf s_src;             // int32  [e_adj]            source rows in (tile, type, row) order
    // windows of k_conv_z (conv_z.cu) over the S format: per 128-row tile the distinct source rows (and the tile's own rows)
    // as contiguous runs; every edge of s_src as a uint16 window row
    bool has_z = false;
    DevBuf zw_meta;           // int32  [s_tiles][4]              {runs, window rows, window row of the tile's first node, edges}
    DevBuf zw_seg;            // int32  [s_tiles][ZW_MAXSEG][2]   {first source row, first window row} per run
    DevBuf z_loc;             // uint16 [e_adj + padding]         window row of s_src[e]
    DevBuf z_tab;             // uint16 [s_passes][128]           per (pass, row): window row of the first source | 0x8000 if several, 0xFFFF if none
    // "T" format for the tcgen05 edge-block kernel (conv_t.cu): super-tiles of t_rows destinations; blocks of 128 same-type
    // slots; quarter q of a block holds destinations with dst % 4 == q, all distinct inside a quarter; every super-tile ends
    // with t_rows / 128 root blocks (type n_types, src = dst = the tile's own rows)
    bool has_t = false;
    int t_rows = 0, t_tiles = 0, t_blocks = 0;
    DevBuf t_bptr;            // int32  [t_tiles + 1]     block range per super-tile
    DevBuf t_btype;           // int32  [t_blocks]        edge type of the block (n_types = root)
    DevBuf t_src;             // int32  [t_blocks * 128]  source row, -1 = empty slot
    DevBuf t_dst;             // uint16 [t_blocks * 128]  destination row inside the super-tile, 0xFFFF = empty slot
    // collision CSR by destination (self loops removed)
    DevBuf col_ptr;           // int32 [n_own + 1]
    DevBuf col_src;           // int32 [e_col]
    // collision windows of k_gin_w (gin_w.cu): per 64-row tile the distinct source rows as contiguous runs
    bool has_gw = false;
    int gw_tiles = 0, gw_direct = 0;   // gw_direct: tiles without a window (sources not local), gathered from global
    DevBuf gw_meta;           // int32  [gw_tiles][8]             {runs (0 = direct), window rows, offset into gw_loc, window row of the tile's first node, edges, -, -, -}
    DevBuf gw_seg;            // int32  [gw_tiles][GW_MAXSEG][2]  {first source row, first window row} per run
    DevBuf gw_loc;            // uint16 [e_col + padding]         window row of every collision edge's source, CSR order, tile blocks 16-byte aligned
    // halo (sharded mode)
    int64_t halo_slot = 0, n_send = 0;
    DevBuf send_rows;         // int32 [n_send]
    DevBuf send_mask;         // uint8 [n_send]: bit q = peer q reads the row (tgnn_set_halo_peers); has_send_mask = false -> all peers
    bool has_send_mask = false;
    DevBuf halo_used;         // uint8 [world * halo_slot]: the mirrored row is read by a local edge
    unsigned need_from = 0xffffffffu;   // peers owning at least one such row
};

struct Scratch {
    std::vector<std::unique_ptr<DevBuf>> bufs;
    size_t next = 0;
    void reset() { next = 0; }
    template <class T> T* get(size_t count) {
        if (next == bufs.size()) bufs.emplace_back(new DevBuf());
        DevBuf& b = *bufs[next++];
        b.reserve(count * sizeof(T) + 16);
        return b.as<T>();
    }
};

// graph_build.cu
void build_graph(Graph& g, Scratch& scratch, int d_e, int64_t n_own, int64_t n_rows,
                 int64_t e_adj, const int64_t* adj_src, const int64_t* adj_dst, const float* adj_feat,
                 int64_t e_col, const int64_t* col_src, const int64_t* col_dst, int want_s /* 0 no, 1 auto, 2 force */,
                 int wn, int want_t /* 0 = no T format, else its super-tile rows: 256 | 512 */, cudaStream_t st);
constexpr int S_BM = 128;         // destination rows per tile of the S format
constexpr int S_OFF_STRIDE = 136; // uint16 per pass (129 used; 272 B keeps 16-byte alignment)
constexpr double S_EDGES_PER_PASS_BREAK_EVEN = 164.0;   // measured: S pass ~5.9 ns, fp16 edge-chunk kernel ~36 ps per edge
constexpr int S_MAX_TYPES = 120;  // the S path is chosen only when K + 1 (root) passes fit its per-tile tables
// k_conv_z: tcgen05 passes with the A operand in tensor memory and the tile's neighbour rows in a shared-memory window
constexpr int ZW_WMAX = 952;      // window rows (119 KB)
constexpr int ZW_MAXSEG = 32;     // contiguous runs per window (one bulk copy each)
constexpr int ZW_CAP = 4608;      // sorted items per tile in the builder = in-edges + own rows
constexpr int ZW_GAP = 2;         // runs closer than this many rows are merged (the gap rows are loaded)
constexpr int ZW_MAX_PASS = 216;  // most edges in one (tile, type) pass
int build_z_windows(Graph& g, Scratch& sc, cudaStream_t st);
int conv_z_blocks(int s_tiles, int sm_count);

// ----- kernels.cu launchers --------------------------------------------------------------------
struct ConvArgs {
    const float* xin;        // [n_rows][32]  b1 of the previous layer (materialised)
    const float* tabF;       // [K+1][2048]   frag tables (hi|lo) of the per-type edge weights; entry K = nnConv.root
    int n_types;             // K
    const float* bias;       // [32]
    const int* cptr; const int* ctype; const int* csrc; const uint8_t* cdst;
    const float* inv_deg;
    float* out;              // pre1 [n_own][32]  LeakyReLU(conv), before BatchNorm
    double* part;            // [n_part][64]  per-CTA partial sums (sum, sum of squares)
    int n_own, n_tiles;
    int wn;                  // rows per warp tile: WN_SMALL or WN_BIG
    int64_t n_chunks;        // 16-edge chunks of the whole graph (launch geometry)
    // fp16-split operands of k_conv_h (conv_h.cu) and the range flags: k_conv_h uses them when both flags are 0 and the
    // 3xTF32 arithmetic on xin (conv_adj_body.cuh) when one is raised
    const uint4* xh;         // [n_rows][8]   split copy of xin: per 4 channels {hi01, hi23, lo01, lo23} fp16 pairs
    const uint32_t* tabH;    // [K+1][1024]   fp16 hi|lo fragment tables; entry K = nnConv.root
    const uint32_t* tabX;    // [K+1][1024]   fp16 hi|lo A-operand fragment tables of W^T (k_conv_x); entry K = nnConv.root
    const int* flag_x;       // raised by the producer of xin when a value is outside the fp16 range
    const int* flag_w;       // raised at table build when a root weight is outside the fp16 range
    const uint8_t* mask;     // node mask (tgnn_set_node_mask) or null: rows with mask 0 are written as 0 and stay out of the statistics
    long long* dbg;          // optional (TGNN_ROLE_DBG=1): per warp of CTA 0 {total, prologue, chunk loop, epilogue} cycles (k_conv_h, one tile per CTA)
};
// Node mask (sub-layout on the resident structures): masked rows are stored as ZERO by every producer, so gathers and
// sums over all neighbours equal sums over the kept ones, and a zero row adds nothing to the BatchNorm sums.
__device__ __forceinline__ bool row_kept(const uint8_t* __restrict__ mask, int node) { return !mask || __ldg(mask + node) != 0; }
// launch geometry shared by k_conv_adj and k_conv_h (same grid => same BatchNorm partial layout):
//   wn = 64 : 8 warps per CTA, 2 CTAs per SM; few tiles (small graphs) => one CTA per tile, k_conv_h splits its chunks
//   wn = 128: 12 warps per CTA (216 KB of accumulator tiles), 1 CTA per SM
struct ConvGeom { int blocks, warps; bool split; int cluster; };   // cluster: CTAs (of one thread-block cluster) sharing a tile in the split geometry
ConvGeom conv_geom(int n_tiles, int wn, int sm_count, int64_t n_chunks);      // n_chunks: 16-edge chunks of the whole graph
inline int conv_adj_num_parts(int n_tiles, int wn, int sm_count, int64_t n_chunks) { ConvGeom g = conv_geom(n_tiles, wn, sm_count, n_chunks); return g.blocks; }   // one partial row per CTA
void launch_conv_adj(const ConvArgs& a, int sm_count, cudaStream_t st);

// fp16-split edge-chunk kernel (conv_h.cu)
constexpr float TG_H_LIMIT = 60000.f;    // |x| above this (or NaN) raises the range flag: fp16 max is 65504
constexpr int TG_HFRAG32 = 1024;         // 32-bit words of one fp16 hi|lo fragment table of a 32x32 matrix
void launch_conv_h(const ConvArgs& a, int sm_count, cudaStream_t st);
void launch_conv_x(const ConvArgs& a, int sm_count, cudaStream_t st);     // transposed MMA roles (weights = A operand), large graphs
constexpr bool TGNN_CONV_X_DEFAULT = false;   // k_conv_x is chosen automatically on large graphs (else only with TGNN_CONV=x)
// tcgen05 edge-block kernel (conv_t.cu) + its fp32 stand-by for the range guard; tabT: [K+1][2048] words, tab32: [K+1][1024]
constexpr int TG_TIMG32 = 2048;          // 32-bit words of one pre-swizzled [64 x 64] fp16 weight image
int conv_t_blocks(int t_tiles, int sm_count);
inline int conv_t_num_parts(int t_tiles, int sm_count) { return conv_t_blocks(t_tiles, sm_count) * 8; }
void launch_conv_t(const ConvArgs& c, const Graph& g, const uint32_t* tabT, const float* tab32, int* err, int sm_count, cudaStream_t st,
                   long long* dbg = nullptr);
// role timing (TGNN_ROLE_DBG=1): acc += cycles spent in expr
#define TGNN_TIMED(acc, expr) ([&]() { const long long _t = clock64(); const bool _r = (expr); (acc) += clock64() - _t; return _r; })()
// tcgen05 "S" formulation of the adjacency branch (conv_s.cu); tabS: [K+1][hi|lo][32][32] transposed weights
void launch_conv_s(const ConvArgs& c, const Graph& g, const float* tabS, int* error_flag, int sm_count, cudaStream_t st);
// tcgen05 passes, A in tensor memory, windowed rows (conv_z.cu); same tabS images
void launch_conv_z(const ConvArgs& c, const Graph& g, const float* tabS, const uint32_t* tabT, int* flag_z, bool force32,
                   int* error_flag, int sm_count, cudaStream_t st, long long* dbg = nullptr);

constexpr int GW_T = 64;          // destination rows per window tile
constexpr int GW_MAXSEG = 32;     // contiguous runs per window (one bulk copy each, one producer lane each)
constexpr int GW_WMAX = 656;      // window rows (82 KB; two buffers per SM)
constexpr int GW_CAP = 3072;      // sorted items per tile in the builder = edges + own rows
constexpr int GW_GAP = 2;         // runs closer than this many rows are merged (the gap rows are loaded)
int build_gin_windows(Graph& g, Scratch& sc, cudaStream_t st);

struct GinArgs {
    const float* xin;        // [n_rows][32]  output of the previous CollConv (written by k_combine), or h0
    const int* col_ptr; const int* col_src;
    // k_gin_w only (null / 0 -> k_gin): windows built by build_gin_windows, and the mapped error word for timeouts
    const uint8_t* mask = nullptr;   // node mask or null
    long long* dbg = nullptr;        // optional: per-warp {cycles, wait 0, wait 1, wait 2} of CTA 0 (TGNN_ROLE_DBG=1)
    const int* gw_meta = nullptr; const int* gw_seg = nullptr; const uint16_t* gw_loc = nullptr; int gw_tiles = 0; int* err = nullptr;
    const float* wfrag;      // 3xTF32 frag tables W1[2048] W2[4096] W3[4096], b1[32] b2[64] b3[32], then fp16 tables W2h[2048] W3h[2048]
    int hmlp;                // layers 2, 3 of the MLP on the fp16 tables (their weights are inside the fp16 range)
    float eps;
    float* out;              // pre2 [n_own][32]
    double* part;            // [n_part][64]
    int n_own;
};
int gin_num_parts(int n_own, int sm_count);
void launch_gin(const GinArgs& a, int sm_count, cudaStream_t st);
int gin_w_blocks(int gw_tiles, int sm_count);
#ifndef TGNN_GW_MLP_WARPS
#define TGNN_GW_MLP_WARPS 7
#endif
constexpr int GW_MLP_WARPS = TGNN_GW_MLP_WARPS;  // MLP warps per CTA of k_gin_w = BatchNorm partial rows per CTA (gather warps: 15 - this)
inline int gin_w_num_parts(int gw_tiles, int sm_count) { return gin_w_blocks(gw_tiles, sm_count) * GW_MLP_WARPS; }
void launch_gin_w(const GinArgs& a, int sm_count, cudaStream_t st);

// b1_new = BN(pre1) * BN(pre2) + residual
// xh / flag (optional): fp16-split copy of the result for k_conv_h and its range flag; g2out (optional): BN(pre2).
// A train-mode BatchNorm that the CONSUMER of the normalised tensor finishes in its own prologue (bn_fin.cuh; small graphs:
// no k_bn_finish launch between producer and consumer).  part: [n_part][2 C] doubles (sum[C] | sum of squares[C]).
struct BnFin {
    const double* part = nullptr; int n_part = 0;
    double count = 1.0; const double* count_ptr = nullptr;      // node mask: number of kept nodes (device), overrides count
    const float* gamma = nullptr; const float* beta = nullptr;
    float* coef_out = nullptr;                                   // [4][C] {mean hi, mean lo, gamma / sqrt(var + eps), beta}; CTA 0 writes it
};

// fin (optional, small graphs): BatchNorm partials of the two branches -- the kernel then computes the coefficients in
// its prologue (and block 0 stores them to coef_out) instead of reading coef1 / coef2.
struct CombineFin {
    const double* part[2] = {nullptr, nullptr}; int n_part[2] = {0, 0}; double count = 1.0;
    const double* count_ptr = nullptr;       // node mask: number of kept nodes (device), overrides count
    const float* gamma[2] = {nullptr, nullptr}; const float* beta[2] = {nullptr, nullptr}; float* coef_out[2] = {nullptr, nullptr};
};
void launch_combine(const float* pre1, const float* coef1, const float* pre2, const float* coef2,
                    const float* residual, float* out, uint4* xh, int* flag, float* g2out, int64_t n_own, cudaStream_t st,
                    const CombineFin* fin = nullptr, const uint8_t* mask = nullptr);

// init MLP: mode 0 = stats of layer 0, 1 = stats of layer 1, 2 = write h0
struct InitArgs {
    const float* x; int d_x;
    const float* w0; const float* b0;    // [32][d_x], [32]
    const float* w1t; const float* b1;   // [32][32] k-major, [32]
    const float* coef0; const float* coef1;
    float* out; double* part; int n_own;
    uint32_t* xh; int* flag;             // mode 2: fp16-split copy of h0 and its range flag (optional)
    const uint8_t* mask;                 // node mask or null
    BnFin fin;                           // fin.part != null: mode 1 finishes stage 0's BatchNorm (coef0) itself, mode 2 stage 1's (coef1)
};
int init_num_parts(int n_own, int sm_count);
// The second Linear's 32 x 32 weights ([k][c], k-major) travel as a KERNEL PARAMETER: every FFMA then takes its weight as a
// constant-bank operand (warp-uniform), instead of through 8 broadcast LDS.128 per 32 FFMA that kept the shared-memory pipe
// 80 % busy (ncu: k_init<1> / <2> l1tex 84 % / 79 %)
struct InitW1 { float w[32 * 32]; };
void launch_init(const InitArgs& a, const InitW1& w1, int mode, int sm_count, cudaStream_t st);

// dense stage of the final MLP: out = LeakyReLU( BN_in(A) @ Wt + b ), plus column statistics
struct DenseArgs {
    const float* const* slabs;  // device array of slab pointers when virtual_concat, else nullptr
    const float* a;             // [n][K] when not virtual_concat
    int virtual_concat;         // A = concat of K/32 slabs of [n_rows][32]
    const float* in_coef;       // [4][K] or nullptr
    const float* wt;            // [K][N_out] k-major
    const float* bias;          // [N_out]
    float* out;                 // [n][N_out]
    double* part;               // [row_blocks][2][N_out]
    int n, K, n_out;
    const uint8_t* mask;        // node mask or null
    BnFin fin;                  // fin.part != null: the kernel finishes the input's BatchNorm (C = K) itself and publishes in_coef
};
int dense_row_blocks(int n);
void launch_dense(const DenseArgs& a, cudaStream_t st);

// tcgen05 (tensor-core) version of the dense stage: w_hi / w_lo = TF32 split of the [N_out][K] weight
void launch_weight_image(const float* w, float* img, int n_out, int K, cudaStream_t st);   // pre-swizzled hi|lo slabs
void launch_weight_image_h(const float* w, void* img, int n_out, int K, int* flag, cudaStream_t st);   // fp16 {hi | lo} slab images of w * 2^6
// w_img16 != null: fp16 kernel + 3xTF32 stand-by (range_flag: int, zeroed per forward); null: 3xTF32 only.  Returns launches.
int launch_dense_tc(const DenseArgs& a, const float* w_img, const void* w_img16, int* range_flag, int* error_flag, int sm_count,
                    cudaStream_t st);

void launch_score(const float* a3, const float* coef, const float* w, float b, float* out,
                  int64_t n, cudaStream_t st, const uint8_t* mask = nullptr, const BnFin* fin = nullptr);

// BatchNorm statistics: reduce partials (fixed order, fp64) and turn them into coefficients.
// part layout: [n_part][2*C] (sum[C], sumsq[C]).  sums_out: [2*C] doubles.
void launch_bn_reduce(const double* part, int n_part, int C, double* sums_out, cudaStream_t st);
void launch_bn_coef(const double* sums, double count, const float* gamma, const float* beta,
                    float* coef_out, int C, cudaStream_t st);
// both steps in one launch for one or two BatchNorms of equal width (single-GPU train mode); sums: [n_bn][2*C]
struct BnFinishArgs {
    const double* part[2]; int n_part[2]; int C; double count;
    const float* gamma[2]; const float* beta[2]; float* coef[2];
    double* sums; unsigned* ticket;
    const double* count_ptr;    // node mask: number of kept nodes (device), overrides count; else null
};
void launch_bn_finish(const BnFinishArgs& a, int n_bn, cudaStream_t st);
// eval mode: coefficients from running statistics
void launch_bn_coef_eval(const float* rmean, const float* rvar, const float* gamma, const float* beta,
                         float* coef_out, int C, cudaStream_t st);

// per-type edge weight tables of all layers in one launch (tables.cu); null table pointers are skipped
struct TableLayer { const float *a1, *c1, *a2, *c2, *a3, *c3, *root; };   // edge MLP (weight, bias) x 3 and nnConv.root
void launch_edge_tables(const float* type_rows, int n_types, int d_e, int n_layers, const TableLayer* layers_dev,
                        float* tabF, float* tabS, uint32_t* tabH, uint32_t* tabT, float* tab32, int* wflags, cudaStream_t st,
                        uint32_t* tabX = nullptr);

void launch_transpose(const float* in, float* out, int rows, int cols, cudaStream_t st);  // out[c][r] = in[r][c]
// frag table (tensor-core B fragments, hi|lo TF32 split) of a k-major [K][N] matrix; maps: see kernels.cu
enum { TG_KMAP_GATHER = 0, TG_KMAP_NATURAL = 1, TG_KMAP_CHAIN = 2, TG_NMAP_NATURAL = 0, TG_NMAP_CONTIG8 = 1 };
constexpr int TG_FRAG32 = 2048;          // floats of a 32x32 frag table
constexpr int TG_GIN_WFLOATS = 2048 + 4096 + 4096 + 128 + 2048 + 2048;
void launch_frag_pack(const float* w_kn, int K, int N, int kmap, int nmap, float* out, cudaStream_t st);
// fp16 hi|lo table in natural k order (GIN layers 2, 3); *flag is raised when a weight is outside the fp16 range
void launch_frag_pack_h16(const float* w_kn, int K, int N, int nmap, float* out, int* flag, cudaStream_t st);

// ---- peer-memory exchange (sharded mode, one process per GPU): every rank maps every peer's exchange buffer with
// CUDA IPC; boundary rows and BatchNorm sums are then WRITTEN straight into the peers' buffers over NVLink by the
// producing kernel and signalled with epoch flags -- no NCCL call, no extra launch, rank-ordered (deterministic) sums.
// Buffer layout (per rank): [flags 4 KB][BN sums 2 x 8 x 512 doubles][halo rows 2 x world x halo_slot x 64 floats];
// everything is double buffered by epoch parity (a rank can run at most one exchange ahead of a peer).
constexpr int PX_MAX_WORLD = 8;
constexpr size_t PX_FLAG_BYTES = 4096;                          // uint32 halo_flag[2][8] at 0, bn_flag[2][8] at 256
constexpr size_t PX_BN_SLOT = 512;                              // doubles per (parity, rank)
constexpr size_t PX_BN_BYTES = 2 * PX_MAX_WORLD * PX_BN_SLOT * sizeof(double);
constexpr size_t PX_HALO_OFF = PX_FLAG_BYTES + PX_BN_BYTES;
struct PeerPtrs { char* base[PX_MAX_WORLD]; int world, rank; int* err; };   // err: mapped host error word
enum { TGNN_DEVERR_PIPELINE = 1, TGNN_DEVERR_PEER = 2 };
// k_bn_finish with the cross-rank sum inside: the last block pushes the local sums to all peers, waits for theirs
void launch_bn_finish_x(const BnFinishArgs& a, int n_bn, const PeerPtrs& p, unsigned epoch, cudaStream_t st);
// boundary rows (a | b) -> slot `rank` of every peer's halo buffer, then the epoch flags (last block)
void launch_halo_push(const float* a, const float* b, const int* rows, int n_send, int64_t halo_slot, const PeerPtrs& p,
                      unsigned epoch, unsigned* ticket, const uint8_t* send_mask /* per row: peers that read it, or null = all */,
                      cudaStream_t st);
// waits for the peers' flags of `epoch`, then unpacks this rank's halo buffer (as launch_halo_unpack)
void launch_halo_unpack_x(const PeerPtrs& p, unsigned epoch, int64_t halo_slot, int64_t n_own, float* a, float* b,
                          uint4* xh, int* flag, const uint8_t* used /* per mirrored row, or null = all */, unsigned need_from,
                          cudaStream_t st);
void launch_mark_halo(const int64_t* src, int64_t e, int64_t n_own, int64_t n_rows, uint8_t* used, cudaStream_t st);

// node mask: kept in-degrees -> inv_deg_masked[n_own]; counters3 = {kept nodes, adjacency edges, collision edges with both
// endpoints kept}; *count = kept nodes as a double (the BatchNorm population)
void launch_node_mask(const Graph& g, const uint8_t* keep, float* inv_deg_masked, int* counters3, double* count, cudaStream_t st);

// halo pack / unpack (sharded mode)
void launch_halo_pack(const float* a, const float* b, const int* rows, int n_send, float* sendbuf, cudaStream_t st);
void launch_halo_unpack(const float* recv, int world, int rank, int64_t halo_slot, int64_t n_own,
                        float* a, float* b, uint4* xh, int* flag, cudaStream_t st);

}  // namespace tgnn
