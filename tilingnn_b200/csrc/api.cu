// C-ABI entry points (include/tgnn.h), parameter store, workspace and the forward orchestration.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "tgnn_internal.h"

// ---- NCCL, bound at run time from the copy torch already loaded (no link-time dependency, so the
// ---- library also loads on machines without NCCL / without a GPU) --------------------------------
namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0 };
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    });
    return api;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != 0) {
        const char* s = nccl().GetErrorString ? nccl().GetErrorString(r) : "?";
        throw tgnn::Error(std::string(what) + ": NCCL error: " + s);
    }
}
std::string g_create_error;
}  // namespace

namespace tgnn {

struct Param {
    std::vector<int64_t> shape;
    DevBuf buf;
    bool set = false;
    size_t numel() const { size_t n = 1; for (auto d : shape) n *= (size_t)d; return n; }
};

struct ProfEntry { std::string fam; cudaEvent_t a, b; };

}  // namespace tgnn

using namespace tgnn;

struct tgnn_handle {
    tgnn_cfg cfg{};
    int sm_count = 148;
    std::string err;
    std::map<std::string, Param> params;           // canonical reference keys
    std::vector<std::string> key_order;
    bool params_dirty = true, tables_dirty = true, graph_set = false;
    // per-type weight tables survive a tgnn_set_graph whose graph has the SAME distinct edge-feature rows (in the builder's
    // order) and picks the same kernels: successive layouts of one tile set share their rows, and the table build
    // (the edge MLP in fp64 for every type and layer, ~0.2 ms at 42 types x 20 layers) is 15 % of a `predict` at N ~ 600
    std::vector<float> type_rows_host;
    uint64_t tables_sig = ~0ull;
    Graph g;
    Scratch scratch;

    // derived parameter layouts
    DevBuf init_w1t;
    InitW1 init_w1_host{};                          // host copy of init_w1t: passed to k_init as a kernel parameter
    std::vector<std::unique_ptr<DevBuf>> gin_wt;    // per layer: frag tables W1|W2|W3 and biases b1|b2|b3
    std::vector<std::unique_ptr<DevBuf>> fin_wt;    // 4: k-major transposes (CUDA-core path)
    std::vector<std::unique_ptr<DevBuf>> fin_whl;   // 4: pre-swizzled hi|lo slab images of the weights (tcgen05 path)
    std::vector<std::unique_ptr<DevBuf>> fin_whh;   // 4: fp16 {hi | lo} slab images of the weights * 2^6 (k_dense_tc<N, true>)
    bool fin_h16[4] = {false, false, false, false}; // the stage's weights fit the fp16 range (checked at pack time)
    bool dense_tf32 = false;                        // TGNN_DENSE=tf32 keeps the final MLP on 3xTF32 (A/B, tests)
    DevBuf dev_error;                               // int[2]: [1] = scratch range flag of pack_params ([0] unused)
    // Device-side error word in MAPPED PINNED HOST memory: kernels store a code there (1 = tcgen05 pipeline timeout,
    // 2 = peer-exchange wait timeout) and the host reads it without a CUDA call -- at the start of every API call, after
    // the synchronous checks, and from tgnn_check_error.  It stays readable even after a kernel trap killed the context.
    int* err_host = nullptr;
    int* err_dev = nullptr;
    bool dense_ffma = false;                        // TGNN_DENSE=ffma selects the CUDA-core dense stage (debug A/B)
    // parameter pointers resolved once per pack_params (no string building / map lookups per launch in the forward)
    struct LayerP { const float *conv_bias, *bn_a_w, *bn_a_b, *bn_c_w, *bn_c_b; };
    struct BnP { const float *w, *b; };
    std::vector<LayerP> lp;
    BnP init_bn[2]{}, fin_bn[4]{};
    const float *init_w0 = nullptr, *init_b0 = nullptr, *init_b1 = nullptr, *fin_bias[4]{}, *score_w = nullptr;
    bool eval_coefs_valid = false;                  // eval-mode coefficients depend on the parameters only
    std::vector<float> gin_eps;
    std::vector<int> gin_hmlp;                      // per layer: GIN MLP layers 2, 3 may use the fp16 tables
    bool gin_tf32_only = getenv("TGNN_GIN") && std::string(getenv("TGNN_GIN")) == "tf32";   // A/B runs
    float fin_last_bias = 0.f;
    DevBuf coef;                                    // all BatchNorm coefficient blocks
    size_t coef_init[2]{}, coef_fin[4]{};
    std::vector<size_t> coef_a, coef_c;
    DevBuf tab;                                     // [L][K+1][2048] frag tables (entry K = root)
    DevBuf tabS;                                    // [L][K+1][2048] transposed hi|lo tables of the tcgen05 conv kernel
    DevBuf table_layers;                            // TableLayer[L]: parameter pointers for tables.cu
    DevBuf tabH;                                    // [L][K+1][1024] fp16 hi|lo fragment tables of k_conv_h
    DevBuf hflags;                                  // int [L+1] activation range flags (per forward) | [L] weight range flags
    bool conv_chunk_only = false;                   // TGNN_CONV=chunk forces the 3xTF32 mma.sync edge-chunk kernel
    bool conv_s_only = false;                       // TGNN_CONV=s forces the tcgen05 S kernel whenever its format exists
    bool conv_h_only = false;                       // TGNN_CONV=h forces the fp16-split edge-chunk kernel (never S)
    bool conv_t_only = false;                       // TGNN_CONV=t forces the tcgen05 edge-block kernel (any graph size, 256-row super-tiles)
    bool conv_z_only = false;                       // TGNN_CONV=z forces the windowed tcgen05 kernel whenever every tile gets a window
    bool use_s = false, use_h = false, use_t = false, use_z = false;   // decided per graph in set_graph
    bool need_xh() const { return use_h || use_t || use_z; }  // the fp16-split copy of b1 is an operand of these kernels
    bool conv_z32 = false;                          // TGNN_CONV=z32: k_conv_z's tf32 variant takes every layer
    int conv_x_mode = -1;                           // TGNN_CONV=x forces the transposed-roles variant of k_conv_h on large graphs, =h forbids it; -1 auto
    bool use_x = false;                             // k_conv_x instead of k_conv_h (large graphs: persistent geometry, 64-row tiles)
    DevBuf tabX;                                    // [L][K+1][1024] fp16 hi|lo A-operand fragment tables of W^T (k_conv_x)
    DevBuf tabT, tab32;                             // [L][K+1] pre-swizzled fp16 weight images / plain fp32 tables of k_conv_t (+ stand-by)
    int tile_rows_forced = 0;                       // TGNN_TILE=64|128 (A/B runs)
    bool tables_streamed = false;                   // many edge types: one layer's weight tables at a time
    int ginw_mode = getenv("TGNN_GINW") ? atoi(getenv("TGNN_GINW")) : -1;   // -1 auto, 0 never, 1 whenever the windows exist (A/B, tests)
    bool use_gw = false;                            // k_gin_w (staged neighbour windows) for this graph

    // workspace
    std::vector<std::unique_ptr<DevBuf>> mid;
    DevBuf pre1, pre2[2], fa[4], partA, partB, sums, slab_ptrs, halo;
    DevBuf xh;                                      // [n_rows][8] uint4: fp16-split copy of the current layer's b1
    int* rflag(int i) { return hflags.as<int>() + i; }
    int* wflag(int i) { return hflags.as<int>() + cfg.depth + 1 + i; }
    int* zflag(int i) { return hflags.as<int>() + 2 * cfg.depth + 3 + i; }
    int* dflag(int k) { return hflags.as<int>() + 3 * cfg.depth + 3 + k; }   // final MLP stage k: an input left the fp16 range (per forward)   // k_conv_z: a multi-edge sum left the fp16 range (per forward)
    unsigned* bn_ticket() { return reinterpret_cast<unsigned*>(hflags.as<int>() + 2 * cfg.depth + 1); }   // k_bn_finish; +1: k_halo_push
    size_t workspace_bytes = 0;

    // sharding
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;

    // peer-memory exchange (CUDA IPC over NVLink); falls back to NCCL collectives when it cannot be set up
    struct PeerX {
        DevBuf buf;                                  // this rank's exchange buffer (tgnn_internal.h: PX_* layout)
        DevBuf staging;                              // handle all-gather staging
        void* peer[PX_MAX_WORLD] = {};               // mapped peer buffers (own slot = buf.p)
        char handles[PX_MAX_WORLD][64] = {};         // the IPC handles currently mapped
        bool ok = false, disabled = getenv("TGNN_P2P") && std::string(getenv("TGNN_P2P")) == "0";
        unsigned epoch_halo = 0, epoch_bn = 0;
        int64_t halo_slot = -1;
    } px;

    // One-submit forward (CUDA graph replay).  The launch sequence of a forward depends only on the resident graph
    // structures, the parameters and the BatchNorm mode -- not on x.  The first forward with a given key runs eagerly
    // (one-shot topologies, as in the reference's greedy loop, never pay for a capture), the second is captured on a
    // private stream with x / scores redirected to staging buffers, later ones are memcpy + cudaGraphLaunch + memcpy
    // on the caller's stream: ~75 launches of a 600-node, 20-layer forward become three submissions.
    struct Replay {
        cudaGraphExec_t exec = nullptr;
        cudaStream_t cap_stream = nullptr;
        uint64_t key = 0; int seen = 0;
        DevBuf x_stage, s_stage;
        int64_t launches = 0;
        bool disabled = getenv("TGNN_GRAPH") && std::string(getenv("TGNN_GRAPH")) == "0";
        void drop() { if (exec) { cudaGraphExecDestroy(exec); exec = nullptr; } seen = 0; }
    } replay;
    uint64_t graph_gen = 0, param_gen = 0;
    // small graphs: a layer's collision branch (k_gin) runs on a side stream next to the adjacency branch (k_conv_*): the two
    // kernels are latency chains of a few CTAs each (16 + 19 us at N ~ 600), not throughput work.  Fork / join by events, so a
    // captured forward gets two parallel branches per layer.
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool two_streams_off = getenv("TGNN_BRANCHES") && std::string(getenv("TGNN_BRANCHES")) == "0";
    bool fin_off = getenv("TGNN_BNFIN") && std::string(getenv("TGNN_BNFIN")) == "launch";       // A/B: BatchNorm statistics by k_bn_finish launches

    // node mask (tgnn_set_node_mask): sub-layout on the resident structures
    bool mask_on = false;
    DevBuf mask_buf, inv_deg_m, mask_cnt;           // uint8 [n_own] | float [n_own] kept in-degree reciprocals | {int[3] counters, pad, double kept nodes}
    const uint8_t* mask() const { return mask_on ? mask_buf.as<uint8_t>() : nullptr; }
    const double* count_ptr() const { return mask_on ? reinterpret_cast<const double*>(mask_cnt.as<char>() + 16) : nullptr; }
    DevBuf role_dbg;                                // TGNN_ROLE_DBG=1: [2 kernels][32 warps][4] cycle counters of CTA 0 (k_conv_t, k_gin_w)
    bool role_dbg_on = getenv("TGNN_ROLE_DBG") != nullptr;

    // bookkeeping
    int64_t launches = 0, collectives = 0;
    int stop_layer = -1;
    int last_layer_run = -1;
    bool profiling = false;
    bool check_errors = getenv("TGNN_CHECK") != nullptr;     // synchronous device-error check after every forward
    std::vector<ProfEntry> prof;
    std::map<std::string, std::pair<float, int>> prof_result;

    bool train_mode_forward() const { return cfg.bn_mode == TGNN_BN_TRAIN; }
    float* P(const std::string& k) {
        auto it = params.find(k);
        TGNN_CHECK(it != params.end(), "internal: unknown parameter " + k);
        return it->second.buf.as<float>();
    }
    float* C(size_t off) { return coef.as<float>() + off; }
};

namespace {

std::string canonical_key(const std::string& k) {
    // graph_networks/layers/edge_conv.py:17-18 registers self.mlp also as self.nnConv.nn
    const std::string alias = ".nnConv.nn.mlp.";
    size_t p = k.find(alias);
    if (p == std::string::npos) return k;
    return k.substr(0, p) + ".mlp.mlp." + k.substr(p + alias.size());
}

void add_param(tgnn_handle* h, const std::string& k, std::vector<int64_t> shape) {
    h->params[k].shape = std::move(shape);
    h->key_order.push_back(k);
}
void add_lt(tgnn_handle* h, const std::string& p, int in, int out, bool bn) {
    add_param(h, p + ".linear.weight", {out, in});
    add_param(h, p + ".linear.bias", {out});
    if (bn) {
        for (const char* s : {".batch_norm.weight", ".batch_norm.bias", ".batch_norm.running_mean", ".batch_norm.running_var"})
            add_param(h, p + s, {out});
    }
}
void add_bn(tgnn_handle* h, const std::string& p, int c) {
    for (const char* s : {".weight", ".bias", ".running_mean", ".running_var"}) add_param(h, p + s, {c});
}

void declare_params(tgnn_handle* h) {
    const int L = h->cfg.depth, dx = h->cfg.d_x, de = h->cfg.d_e;
    add_lt(h, "init_node_feature_trans.mlp.0", dx, F, true);
    add_lt(h, "init_node_feature_trans.mlp.1", F, F, true);
    for (int i = 0; i < L; ++i) {
        std::string p = "brch_1_graph_conv_layers." + std::to_string(i);
        int dims[4] = {de, 32, 64, F * F};
        for (int k = 0; k < 3; ++k) add_lt(h, p + ".mlp.mlp." + std::to_string(k), dims[k], dims[k + 1], false);
        add_param(h, p + ".nnConv.root", {F, F});
        add_param(h, p + ".nnConv.bias", {F});
        add_bn(h, p + ".batch_norm", F);
    }
    for (int i = 0; i < L; ++i) {
        std::string p = "brch_2_coll_conv_layers." + std::to_string(i);
        add_param(h, p + ".ginConv.eps", {1});
        int dims[4] = {F, 32, 64, F};
        for (int k = 0; k < 3; ++k) add_lt(h, p + ".ginConv.nn.mlp." + std::to_string(k), dims[k], dims[k + 1], false);
        add_bn(h, p + ".batch_norm", F);
    }
    int dims[5] = {F * (L + 1), 256, 128, 64, F};
    for (int k = 0; k < 4; ++k) add_lt(h, "final_mlp.0.mlp." + std::to_string(k), dims[k], dims[k + 1], true);
    add_lt(h, "final_mlp.1", F, 1, false);
}

bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

const int FIN_DIMS[5] = {0, 256, 128, 64, 32};

void pack_params(tgnn_handle* h, cudaStream_t st) {
    if (!h->params_dirty) return;
    for (auto& k : h->key_order) TGNN_CHECK(h->params[k].set, "parameter not set: " + k);
    const int L = h->cfg.depth;
    const char* dsel = getenv("TGNN_DENSE");
    h->dense_ffma = dsel && std::string(dsel) == "ffma";
    h->dense_tf32 = dsel && std::string(dsel) == "tf32";
    h->init_w1t.reserve(32 * 32 * sizeof(float));
    launch_transpose(h->P("init_node_feature_trans.mlp.1.linear.weight"), h->init_w1t.as<float>(), 32, 32, st);
    TGNN_CUDA(cudaMemcpyAsync(h->init_w1_host.w, h->init_w1t.p, sizeof(h->init_w1_host.w), cudaMemcpyDeviceToHost, st));
    TGNN_CUDA(cudaStreamSynchronize(st));        // (parameters change rarely; the launches below read the host copy)
    h->gin_wt.clear(); h->gin_eps.assign(L, 0.f); h->gin_hmlp.assign(L, 0);
    DevBuf tmp_t;
    for (int i = 0; i < L; ++i) {
        std::string p = "brch_2_coll_conv_layers." + std::to_string(i);
        int dims[4] = {F, 32, 64, F};
        h->gin_wt.emplace_back(new DevBuf());
        h->gin_wt.back()->reserve((size_t)TG_GIN_WFLOATS * sizeof(float));
        float* wf = h->gin_wt.back()->as<float>();
        const size_t off[3] = {0, 2048, 2048 + 4096};
        const int kmap[3] = {TG_KMAP_NATURAL, TG_KMAP_CHAIN, TG_KMAP_CHAIN};
        const int nmap[3] = {TG_NMAP_NATURAL, TG_NMAP_NATURAL, TG_NMAP_CONTIG8};
        float* boff = wf + 2048 + 4096 + 4096;
        for (int k = 0; k < 3; ++k) {
            const std::string lin = p + ".ginConv.nn.mlp." + std::to_string(k) + ".linear";
            tmp_t.reserve((size_t)dims[k] * dims[k + 1] * sizeof(float));
            launch_transpose(h->P(lin + ".weight"), tmp_t.as<float>(), dims[k + 1], dims[k], st);       // -> [K][N]
            launch_frag_pack(tmp_t.as<float>(), dims[k], dims[k + 1], kmap[k], nmap[k], wf + off[k], st);
            if (k >= 1)          // fp16 tables of layers 2, 3 behind the 3xTF32 block; range flag in the scratch int
                launch_frag_pack_h16(tmp_t.as<float>(), dims[k], dims[k + 1], nmap[k], wf + 2048 + 4096 + 4096 + 128 + (k - 1) * 2048,
                                     h->dev_error.as<int>() + 1, st);
            TGNN_CUDA(cudaMemcpyAsync(boff, h->P(lin + ".bias"), dims[k + 1] * sizeof(float), cudaMemcpyDeviceToDevice, st));
            boff += dims[k + 1];
            TGNN_CUDA(cudaStreamSynchronize(st));                                                        // tmp_t is reused
        }
        TGNN_CUDA(cudaMemcpyAsync(&h->gin_eps[i], h->P(p + ".ginConv.eps"), sizeof(float), cudaMemcpyDeviceToHost, st));
        int out_of_range = 0;
        TGNN_CUDA(cudaMemcpyAsync(&out_of_range, h->dev_error.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        TGNN_CUDA(cudaStreamSynchronize(st));
        h->gin_hmlp[i] = out_of_range == 0 && !h->gin_tf32_only;
        TGNN_CUDA(cudaMemsetAsync(h->dev_error.as<int>() + 1, 0, sizeof(int), st));
    }
    h->fin_wt.clear(); h->fin_whl.clear(); h->fin_whh.clear();
    int dims[5] = {F * (L + 1), 256, 128, 64, F};
    for (int k = 0; k < 4; ++k) {
        const float* w = h->P("final_mlp.0.mlp." + std::to_string(k) + ".linear.weight");
        const size_t ne = (size_t)dims[k] * dims[k + 1];
        h->fin_wt.emplace_back(new DevBuf());
        h->fin_wt.back()->reserve(ne * sizeof(float));
        launch_transpose(w, h->fin_wt.back()->as<float>(), dims[k + 1], dims[k], st);
        h->fin_whl.emplace_back(new DevBuf());
        h->fin_whl.back()->reserve(2 * ne * sizeof(float));
        launch_weight_image(w, h->fin_whl.back()->as<float>(), dims[k + 1], dims[k], st);
        h->fin_whh.emplace_back(new DevBuf());
        h->fin_whh.back()->reserve(2 * ne * sizeof(uint16_t));
        TGNN_CUDA(cudaMemsetAsync(h->dev_error.as<int>() + 1, 0, sizeof(int), st));
        launch_weight_image_h(w, h->fin_whh.back()->p, dims[k + 1], dims[k], h->dev_error.as<int>() + 1, st);
        int out_of_range = 0;
        TGNN_CUDA(cudaMemcpyAsync(&out_of_range, h->dev_error.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        TGNN_CUDA(cudaStreamSynchronize(st));
        h->fin_h16[k] = out_of_range == 0 && !h->dense_tf32;
        TGNN_CUDA(cudaMemsetAsync(h->dev_error.as<int>() + 1, 0, sizeof(int), st));
    }
    {
        std::vector<TableLayer> tl(L);
        for (int i = 0; i < L; ++i) {
            const std::string c = "brch_1_graph_conv_layers." + std::to_string(i), p = c + ".mlp.mlp.";
            tl[i] = TableLayer{h->P(p + "0.linear.weight"), h->P(p + "0.linear.bias"), h->P(p + "1.linear.weight"),
                               h->P(p + "1.linear.bias"), h->P(p + "2.linear.weight"), h->P(p + "2.linear.bias"),
                               h->P(c + ".nnConv.root")};
        }
        h->table_layers.reserve(L * sizeof(TableLayer));
        TGNN_CUDA(cudaMemcpyAsync(h->table_layers.p, tl.data(), L * sizeof(TableLayer), cudaMemcpyHostToDevice, st));
        TGNN_CUDA(cudaStreamSynchronize(st));
    }
    h->lp.resize(L);
    for (int i = 0; i < L; ++i) {
        const std::string a = "brch_1_graph_conv_layers." + std::to_string(i), c = "brch_2_coll_conv_layers." + std::to_string(i);
        h->lp[i] = {h->P(a + ".nnConv.bias"), h->P(a + ".batch_norm.weight"), h->P(a + ".batch_norm.bias"),
                    h->P(c + ".batch_norm.weight"), h->P(c + ".batch_norm.bias")};
    }
    for (int k = 0; k < 2; ++k) {
        const std::string b = "init_node_feature_trans.mlp." + std::to_string(k) + ".batch_norm";
        h->init_bn[k] = {h->P(b + ".weight"), h->P(b + ".bias")};
    }
    for (int k = 0; k < 4; ++k) {
        const std::string b = "final_mlp.0.mlp." + std::to_string(k);
        h->fin_bn[k] = {h->P(b + ".batch_norm.weight"), h->P(b + ".batch_norm.bias")};
        h->fin_bias[k] = h->P(b + ".linear.bias");
    }
    h->init_w0 = h->P("init_node_feature_trans.mlp.0.linear.weight"); h->init_b0 = h->P("init_node_feature_trans.mlp.0.linear.bias");
    h->init_b1 = h->P("init_node_feature_trans.mlp.1.linear.bias"); h->score_w = h->P("final_mlp.1.linear.weight");
    h->eval_coefs_valid = false;
    TGNN_CUDA(cudaMemcpyAsync(&h->fin_last_bias, h->P("final_mlp.1.linear.bias"), sizeof(float), cudaMemcpyDeviceToHost, st));
    // coefficient blocks
    size_t off = 0;
    auto take = [&](int c) { size_t o = off; off += 4 * (size_t)c; return o; };
    h->coef_init[0] = take(32); h->coef_init[1] = take(32);
    h->coef_a.resize(L); h->coef_c.resize(L);
    for (int i = 0; i < L; ++i) { h->coef_a[i] = take(32); h->coef_c[i] = take(32); }
    for (int k = 0; k < 4; ++k) h->coef_fin[k] = take(FIN_DIMS[k + 1]);
    h->coef.reserve(off * sizeof(float));
    TGNN_CUDA(cudaStreamSynchronize(st));
    h->params_dirty = false;
    h->tables_dirty = true;
}

void eval_coefs(tgnn_handle* h, cudaStream_t st) {
    auto one = [&](const std::string& bn, size_t off, int c) {
        launch_bn_coef_eval(h->P(bn + ".running_mean"), h->P(bn + ".running_var"), h->P(bn + ".weight"), h->P(bn + ".bias"),
                            h->C(off), c, st);
    };
    const int L = h->cfg.depth;
    one("init_node_feature_trans.mlp.0.batch_norm", h->coef_init[0], 32);
    one("init_node_feature_trans.mlp.1.batch_norm", h->coef_init[1], 32);
    for (int i = 0; i < L; ++i) {
        one("brch_1_graph_conv_layers." + std::to_string(i) + ".batch_norm", h->coef_a[i], 32);
        one("brch_2_coll_conv_layers." + std::to_string(i) + ".batch_norm", h->coef_c[i], 32);
    }
    for (int k = 0; k < 4; ++k) one("final_mlp.0.mlp." + std::to_string(k) + ".batch_norm", h->coef_fin[k], FIN_DIMS[k + 1]);
}

// Few edge types (the shipped tile sets: 20-41; the synthetic configs: <= 51): the tables of all layers are built once
// per graph.  Many types (continuous edge features -> up to one type per edge): one layer's tables at a time, built
// right before that layer runs ("streamed"), so memory stays at 12 KB per type instead of 12 KB x depth.
constexpr size_t TABLES_RESIDENT_BYTES = size_t(1) << 30;

void build_tables(tgnn_handle* h, cudaStream_t st, int layer = -1) {
    const int L = h->cfg.depth, K = h->g.n_types;
    const size_t per_layer = (size_t)(K + 1);
    h->tables_streamed = per_layer * L * (TG_FRAG32 * sizeof(float) + TG_HFRAG32 * sizeof(uint32_t)) > TABLES_RESIDENT_BYTES;
    if (layer < 0 && (h->tables_streamed || !h->tables_dirty)) return;
    if (layer >= 0 && !h->tables_streamed) return;
    const int nl = layer < 0 ? L : 1, l0 = layer < 0 ? 0 : layer;
    const size_t slots = per_layer * nl;
    // 3xTF32 fragments: always (k_conv_adj is also the wide-range stand-in of k_conv_h)
    h->tab.reserve(slots * TG_FRAG32 * sizeof(float));
    if (h->need_xh()) TGNN_CUDA(cudaMemsetAsync(h->wflag(l0), 0, (size_t)nl * sizeof(int), st));
    if (h->use_h) h->tabH.reserve(slots * TG_HFRAG32 * sizeof(uint32_t));
    if (h->use_x) h->tabX.reserve(slots * TG_HFRAG32 * sizeof(uint32_t));
    if (h->use_t || h->use_z) h->tabT.reserve(slots * TG_TIMG32 * sizeof(uint32_t));
    if (h->use_t) h->tab32.reserve(slots * F * F * sizeof(float));
    if (h->use_s || h->use_z) h->tabS.reserve(slots * TG_FRAG32 * sizeof(float));
    launch_edge_tables(h->g.type_rows.as<float>(), K, h->cfg.d_e, nl, h->table_layers.as<TableLayer>() + l0, h->tab.as<float>(),
                       (h->use_s || h->use_z) ? h->tabS.as<float>() : nullptr, h->use_h ? h->tabH.as<uint32_t>() : nullptr,
                       (h->use_t || h->use_z) ? h->tabT.as<uint32_t>() : nullptr, h->use_t ? h->tab32.as<float>() : nullptr,
                       h->need_xh() ? h->wflag(l0) : nullptr, st, h->use_x ? h->tabX.as<uint32_t>() : nullptr);
    if (layer < 0) h->tables_dirty = false;
}

// Cost model from B200 measurements (1M nodes, deg 32, 51 types): the edge-chunk mma.sync kernel costs ~72 ps per
// adjacency edge, the tcgen05 S kernel ~5.9 ns per (128-row tile, edge type) pass -> S pays off when the tiles see
// few types relative to their edge count (the shipped tile graphs: 20-41 types), chunk when types are many.
// The windowed tcgen05 kernel (conv_z.cu) works on the same S format.  It is OPT-IN (TGNN_CONV=z | z32): measured on B200 at
// 1M nodes x deg 32 (profiles/r2/conv_z_*.txt) it runs at 1.5 ms per launch against the 1.2 ms of k_conv_h -- its MMA side needs
// only ~180 cycles per (tile, type) pass, but the gather warps spend ~430 instructions per pass and warp (96 selects to undo the
// bank-conflict-free rotated row reads, ~150 for the multi-edge rows of the synthetic graphs) and are latency/issue bound.
void choose_conv_kernel(tgnn_handle* h, cudaStream_t st) {
    h->use_z = false; h->g.has_z = false;
    if (h->g.s_built && h->conv_z_only && h->g.s_max_pass <= ZW_MAX_PASS) {
        h->g.has_z = build_z_windows(h->g, h->scratch, st) == 0;
        h->use_z = h->g.has_z;
    }
    h->use_s = !h->use_z && h->g.has_s && !h->conv_h_only && !h->conv_z_only &&
               (h->conv_s_only || (double)h->g.s_passes * S_EDGES_PER_PASS_BREAK_EVEN < (double)h->g.e_adj);
    h->use_t = h->g.has_t && !h->conv_h_only && !h->conv_s_only && !h->conv_chunk_only && !h->use_z;
    if (h->use_t) h->use_s = false;
    h->use_h = !h->use_s && !h->use_t && !h->use_z && !h->conv_chunk_only;
    // transposed MMA roles (no register moves to assemble the A operand): the persistent large-graph geometry only
    h->use_x = h->use_h && h->conv_x_mode != 0 && h->g.wn == WN_SMALL && !conv_geom(h->g.n_tiles, h->g.wn, h->sm_count, h->g.n_chunks).split &&
               (h->conv_x_mode == 1 || TGNN_CONV_X_DEFAULT);
}

// The tcgen05 edge-block kernel (conv_t.cu) is OPT-IN (TGNN_CONV=t, 256-row super-tiles; TGNN_CONV_T_ROWS=512 for large
// graphs).  Measured on B200 at 1M nodes x deg 32 (profiles/r2): its MMA and epilogue side runs at ~450 cycles per
// 128-edge block, but every way of gathering the blocks' rows (cp.async, TMA gather4, register-staged loads) delivers
// only ~2.6-3.2 TB/s of scattered 128-byte rows out of L2, i.e. 1.3-2.7 ms per launch against the 1.14 ms of k_conv_h,
// whose dst-tile-local walk gets 28 % of its rows from L1.  0 = do not build the format.
int want_t_rows(tgnn_handle* h, int64_t n_own) {
    if (!h->conv_t_only) return 0;
    if (getenv("TGNN_CONV_T_ROWS")) { const int r = atoi(getenv("TGNN_CONV_T_ROWS")); if (r == 256 || r == 512) return r; }
    return n_own >= (int64_t)2 * h->sm_count * 512 ? 512 : 256;
}

// Staged-window collision kernel (gin_w.cu): worth it when the graph is large enough to fill the SMs with 64-row tiles
// and local enough that most tiles get a window; otherwise k_gin's per-lane gathers stay.
constexpr int64_t GINW_MIN_NODES = 32768;
void choose_gin_kernel(tgnn_handle* h, cudaStream_t st) {
    h->use_gw = false; h->g.has_gw = false; h->g.gw_direct = 0;
    if (h->ginw_mode == 0 || h->g.e_col <= 0) return;
    if (h->ginw_mode < 0 && h->g.n_own < GINW_MIN_NODES) return;
    h->g.gw_direct = build_gin_windows(h->g, h->scratch, st);
    h->g.has_gw = true;
    h->use_gw = h->ginw_mode == 1 || 4 * (int64_t)h->g.gw_direct <= (int64_t)h->g.gw_tiles;
}

// 128-row warp tiles give longer same-type runs (half the weight-table reloads, ~10 % fewer padded slots) but only 12
// resident warps per SM instead of 16; measured on B200 at 1M nodes x deg 32 the two cancel (7.5 vs 7.3 ms per forward),
// so 64 stays the default and TGNN_TILE=128 is kept for A/B runs.
int want_s_mode(tgnn_handle* h) { return (h->conv_s_only || h->conv_z_only) ? 2 : ((h->conv_chunk_only || h->conv_h_only || h->conv_t_only) ? 0 : 1); }
int tile_rows_for(tgnn_handle* h, int64_t) { return h->tile_rows_forced ? h->tile_rows_forced : WN_SMALL; }

void alloc_workspace(tgnn_handle* h) {
    const int L = h->cfg.depth;
    const size_t rows = (size_t)h->g.n_rows, own = (size_t)h->g.n_own;
    size_t total = 0;
    auto res = [&](DevBuf& b, size_t bytes) { b.reserve(bytes); total += b.cap; };
    while ((int)h->mid.size() < L + 1) h->mid.emplace_back(new DevBuf());
    for (int i = 0; i <= L; ++i) res(*h->mid[i], rows * F * sizeof(float));
    res(h->pre1, own * F * sizeof(float));
    if (h->need_xh()) res(h->xh, rows * F * sizeof(float));
    res(h->pre2[0], rows * F * sizeof(float));
    res(h->pre2[1], rows * F * sizeof(float));
    for (int k = 0; k < 4; ++k) res(h->fa[k], own * FIN_DIMS[k + 1] * sizeof(float));
    size_t np = std::max({(size_t)conv_adj_num_parts(h->g.n_tiles, h->g.wn, h->sm_count, h->g.n_chunks), (size_t)h->g.s_tiles, (size_t)gin_num_parts((int)own, h->sm_count),
                          (size_t)(h->g.has_t ? conv_t_num_parts(h->g.t_tiles, h->sm_count) : 0),
                          (size_t)gin_w_num_parts(h->g.gw_tiles, h->sm_count),
                          (size_t)init_num_parts((int)own, h->sm_count)});
    size_t part_bytes = std::max(np * 64, (size_t)dense_row_blocks((int)own) * 2 * 256) * sizeof(double);
    res(h->partA, part_bytes);
    res(h->partB, part_bytes);                       // (stages alternate between the two: a consumer that finishes a BatchNorm
                                                     //  itself reads one while its CTAs write the other)
    res(h->sums, 2 * 512 * sizeof(double));
    res(h->slab_ptrs, (size_t)(L + 1) * sizeof(float*));
    std::vector<const float*> slabs(L + 1);
    for (int i = 0; i <= L; ++i) slabs[i] = h->mid[i]->as<float>();
    TGNN_CUDA(cudaMemcpy(h->slab_ptrs.p, slabs.data(), slabs.size() * sizeof(float*), cudaMemcpyHostToDevice));
    if (h->world > 1) {
        const void* before = h->halo.p;
        res(h->halo, (size_t)h->world * h->g.halo_slot * 64 * sizeof(float));
        if (h->halo.p != before) TGNN_CUDA(cudaMemset(h->halo.p, 0, h->halo.cap));       // padding rows are never written
    }
    h->workspace_bytes = total;
}

constexpr int64_t SYNC_CHECK_MAX_NODES = 16384;

struct Launcher {
    tgnn_handle* h; cudaStream_t st;
    void begin(const char* fam) {
        if (!h->profiling) return;
        ProfEntry e; e.fam = fam;
        cudaEventCreate(&e.a); cudaEventCreate(&e.b);
        cudaEventRecord(e.a, st);
        h->prof.push_back(e);
    }
    // n_work: launches that do the family's work (k_conv_h's stand-by twin k_conv_adj exits at once and is not one)
    void end(int n_launch, int n_work = -1) {
        h->launches += n_launch;
        if (!h->profiling) return;
        cudaEventRecord(h->prof.back().b, st);
        h->prof_result[h->prof.back().fam].second += n_work < 0 ? n_launch : n_work;
    }
};

const char* device_error_text(int code) {
    switch (code) {
        case TGNN_DEVERR_PIPELINE: return "device-side pipeline timeout in a tcgen05 kernel (mbarrier wait exceeded its bound); the scores of that forward are invalid";
        case TGNN_DEVERR_PEER: return "peer-memory exchange timed out waiting for another rank (a rank stalled or died); the kernel was aborted";
        default: return "unknown device-side error code";
    }
}
// Host read of the mapped error word; throws (and clears the word) when a kernel reported an error.
void check_device_error(tgnn_handle* h, const char* where) {
    if (!h->err_host) return;
    const int e = *reinterpret_cast<volatile int*>(h->err_host);
    if (e == 0) return;
    *reinterpret_cast<volatile int*>(h->err_host) = 0;
    throw Error(std::string(where) + ": " + device_error_text(e));
}

PeerPtrs peer_ptrs(tgnn_handle* h) {
    PeerPtrs p{};
    p.err = h->err_dev;
    for (int q = 0; q < h->world; ++q) p.base[q] = static_cast<char*>(h->px.peer[q]);
    p.world = h->world; p.rank = h->rank;
    return p;
}

void peer_close(tgnn_handle* h) {
    for (int q = 0; q < PX_MAX_WORLD; ++q) {
        if (h->px.peer[q] && q != h->rank) cudaIpcCloseMemHandle(h->px.peer[q]);
        h->px.peer[q] = nullptr;
        memset(h->px.handles[q], 0, 64);
    }
    h->px.ok = false;
}

// Collective (every rank calls it from tgnn_set_graph_shard): size this rank's exchange buffer for the graph, publish
// its IPC handle through a 64-byte NCCL all-gather, map the peers' buffers.  Any failure on any rank -> all ranks keep
// the NCCL collectives (agreed through an all-reduce).
void peer_setup(tgnn_handle* h, cudaStream_t st) {
    tgnn_handle::PeerX& x = h->px;
    if (h->world <= 1 || h->world > PX_MAX_WORLD || x.disabled) { x.ok = false; return; }
    const size_t need = PX_HALO_OFF + 2 * (size_t)h->world * (size_t)h->g.halo_slot * 64 * sizeof(float);
    const void* before = x.buf.p;
    if (need > x.buf.cap && x.buf.p) {
        // The buffer is exported through CUDA IPC: freeing it while a peer still has it mapped is undefined behaviour.
        // halo_slot is the same on every rank (shard.make_plan takes the maximum), so every rank grows in the SAME
        // call: close the mappings of the peers' buffers first, then a collective barrier, then reallocate.
        peer_close(h);
        x.staging.reserve((size_t)(h->world + 1) * 64 + 2 * sizeof(float));
        float* bar = reinterpret_cast<float*>(x.staging.as<char>());
        TGNN_CUDA(cudaMemsetAsync(bar, 0, 2 * sizeof(float), st));
        nccl_check(nccl().AllReduce(bar, bar + 1, 1, ncclFloat32, ncclSum, h->comm, st), "peer buffer regrow barrier");
        TGNN_CUDA(cudaStreamSynchronize(st));
    }
    x.buf.reserve(need);
    if (x.buf.p != before || x.halo_slot != h->g.halo_slot) {
        TGNN_CUDA(cudaMemsetAsync(x.buf.p, 0, x.buf.cap, st));        // flags, and no garbage in padding rows
        x.epoch_halo = 0; x.epoch_bn = 0;
    }
    x.halo_slot = h->g.halo_slot;
    x.staging.reserve((size_t)(h->world + 1) * 64 + 2 * sizeof(float));
    cudaIpcMemHandle_t mine;
    float good = cudaIpcGetMemHandle(&mine, x.buf.p) == cudaSuccess ? 1.f : 0.f;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    char* stg = x.staging.as<char>();
    TGNN_CUDA(cudaMemcpyAsync(stg, &mine, 64, cudaMemcpyHostToDevice, st));
    nccl_check(nccl().AllGather(stg, stg + 64, 16, ncclFloat32, h->comm, st), "IPC handle all-gather");
    std::vector<char> all((size_t)h->world * 64);
    TGNN_CUDA(cudaMemcpyAsync(all.data(), stg + 64, all.size(), cudaMemcpyDeviceToHost, st));
    TGNN_CUDA(cudaStreamSynchronize(st));                             // (also orders every rank's memset before any push)
    for (int q = 0; q < h->world && good > 0.f; ++q) {
        if (q == h->rank) { x.peer[q] = x.buf.p; continue; }
        if (x.peer[q] && memcmp(x.handles[q], all.data() + (size_t)q * 64, 64) == 0) continue;     // still mapped
        if (x.peer[q]) { cudaIpcCloseMemHandle(x.peer[q]); x.peer[q] = nullptr; }
        cudaIpcMemHandle_t hq;
        memcpy(&hq, all.data() + (size_t)q * 64, 64);
        if (cudaIpcOpenMemHandle(&x.peer[q], hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            x.peer[q] = nullptr; good = 0.f;
        } else {
            memcpy(x.handles[q], &hq, 64);
        }
    }
    float* flagp = reinterpret_cast<float*>(stg + (size_t)(h->world + 1) * 64);
    TGNN_CUDA(cudaMemcpyAsync(flagp, &good, sizeof(float), cudaMemcpyHostToDevice, st));
    nccl_check(nccl().AllReduce(flagp, flagp + 1, 1, ncclFloat32, ncclSum, h->comm, st), "peer setup agreement");
    float total = 0.f;
    TGNN_CUDA(cudaMemcpyAsync(&total, flagp + 1, sizeof(float), cudaMemcpyDeviceToHost, st));
    TGNN_CUDA(cudaStreamSynchronize(st));
    x.ok = total > (float)h->world - 0.5f;
    if (!x.ok) peer_close(h);
}

void allreduce_sums(tgnn_handle* h, double* sums, int n, cudaStream_t st) {
    if (h->world <= 1) return;
    nccl_check(nccl().AllReduce(sums, sums, (size_t)n, ncclFloat64, ncclSum, h->comm, st), "BatchNorm all-reduce");
    h->collectives += 1;
}

void halo_exchange(tgnn_handle* h, float* a, float* b, int* flag, cudaStream_t st, Launcher& lz) {
    if (h->world <= 1) return;
    lz.begin("halo");
    if (h->px.ok) {
        // boundary rows go straight into the peers' buffers (NVLink stores), flags follow; the unpack waits for the flags
        const unsigned epoch = ++h->px.epoch_halo;
        const PeerPtrs pp = peer_ptrs(h);
        launch_halo_push(a, b, h->g.send_rows.as<int>(), (int)h->g.n_send, h->g.halo_slot, pp, epoch, h->bn_ticket() + 1,
                         h->g.has_send_mask ? h->g.send_mask.as<uint8_t>() : nullptr, st);
        launch_halo_unpack_x(pp, epoch, h->g.halo_slot, h->g.n_own, a, b, h->need_xh() ? h->xh.as<uint4>() : nullptr, flag,
                             h->g.halo_used.as<uint8_t>(), h->g.need_from, st);
        h->collectives += 1;
        lz.end(2);
        return;
    }
    float* slot = h->halo.as<float>() + (size_t)h->rank * h->g.halo_slot * 64;
    launch_halo_pack(a, b, h->g.send_rows.as<int>(), (int)h->g.n_send, slot, st);
    nccl_check(nccl().AllGather(slot, h->halo.p, (size_t)h->g.halo_slot * 64, ncclFloat32, h->comm, st), "halo all-gather");
    launch_halo_unpack(h->halo.as<float>(), h->world, h->rank, h->g.halo_slot, h->g.n_own, a, b,
                       h->need_xh() ? h->xh.as<uint4>() : nullptr, flag, st);
    h->collectives += 1;
    lz.end(2);
}

constexpr int64_t FIN_MAX_NODES = 32768;          // up to here the consumers finish the BatchNorm statistics themselves (bn_fin.cuh)
constexpr int64_t BRANCH_MAX_NODES = int64_t(1) << 40;   // measured: -5..-11 % up to 100k nodes, -1 % at 300k, -0.3 % at 1M (the tail of one
                                                        // branch overlaps the head of the other), never slower: always on (TGNN_BRANCH_MAX overrides)

void forward_impl(tgnn_handle* h, const float* x, float* scores, cudaStream_t st, bool capturing = false) {
    TGNN_CHECK(h->graph_set, "tgnn_forward: no graph set (call tgnn_set_graph first)");
    if (!capturing) {
        check_device_error(h, "tgnn_forward (reported by an earlier forward)");
        pack_params(h, st);
        build_tables(h, st);
    }
    const int L = h->cfg.depth;
    const bool train = h->cfg.bn_mode == TGNN_BN_TRAIN;
    const int n_own = (int)h->g.n_own;
    const double count = (double)h->g.n_global;
    h->launches = 0; h->collectives = 0;
    for (auto& e : h->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    h->prof.clear(); h->prof_result.clear();
    Launcher lz{h, st};
    double* sums = h->sums.as<double>();
    if (h->role_dbg_on && !h->role_dbg.p) { h->role_dbg.reserve(256 * sizeof(long long)); TGNN_CUDA(cudaMemset(h->role_dbg.p, 0, 256 * sizeof(long long))); }
    if (!train && !h->eval_coefs_valid) { lz.begin("bnfin"); eval_coefs(h, st); lz.end(2 + 2 * L + 4); h->eval_coefs_valid = true; }
    if (train) h->eval_coefs_valid = false;          // train-mode forwards overwrite the coefficient blocks
    if (h->need_xh()) TGNN_CUDA(cudaMemsetAsync(h->rflag(0), 0, (size_t)(L + 1) * sizeof(int), st));
    if (h->use_z) TGNN_CUDA(cudaMemsetAsync(h->zflag(0), 0, (size_t)L * sizeof(int), st));
    TGNN_CUDA(cudaMemsetAsync(h->dflag(0), 0, 4 * sizeof(int), st));

    auto finish_bn = [&](const double* part, int n_part, int c, const tgnn_handle::BnP& bn, size_t coef_off) {
        if (h->world == 1 || h->px.ok) {
            BnFinishArgs fa{};
            fa.part[0] = part; fa.n_part[0] = n_part; fa.C = c; fa.count = count;
            fa.gamma[0] = bn.w; fa.beta[0] = bn.b; fa.coef[0] = h->C(coef_off);
            fa.sums = sums; fa.ticket = h->bn_ticket(); fa.count_ptr = h->count_ptr();
            if (h->world == 1) launch_bn_finish(fa, 1, st);
            else { launch_bn_finish_x(fa, 1, peer_ptrs(h), ++h->px.epoch_bn, st); h->collectives += 1; }
            return;
        }
        launch_bn_reduce(part, n_part, c, sums, st);
        allreduce_sums(h, sums, 2 * c, st);
        launch_bn_coef(sums, count, bn.w, bn.b, h->C(coef_off), c, st);
    };

    // ---- init MLP ------------------------------------------------------------------------------
    InitArgs ia{};
    ia.x = x; ia.d_x = h->cfg.d_x;
    ia.w0 = h->init_w0; ia.b0 = h->init_b0;
    ia.w1t = h->init_w1t.as<float>(); ia.b1 = h->init_b1;
    ia.coef0 = h->C(h->coef_init[0]); ia.coef1 = h->C(h->coef_init[1]);
    ia.out = h->mid[0]->as<float>(); ia.part = h->partA.as<double>(); ia.n_own = n_own;
    ia.xh = h->need_xh() ? h->xh.as<uint32_t>() : nullptr; ia.flag = h->need_xh() ? h->rflag(0) : nullptr;
    ia.mask = h->mask();
    const int np_init = init_num_parts(n_own, h->sm_count);
    // small graphs: no k_bn_finish launches -- every consumer of a BatchNorm'd tensor finishes the statistic in its prologue
    const bool fin_small = train && h->world == 1 && n_own <= FIN_MAX_NODES && !h->fin_off;
    auto make_fin = [&](const double* part, int n_part, const tgnn_handle::BnP& bn, size_t coef_off) {
        BnFin f{};
        f.part = part; f.n_part = n_part; f.count = count; f.count_ptr = h->count_ptr();
        f.gamma = bn.w; f.beta = bn.b; f.coef_out = h->C(coef_off);
        return f;
    };
    if (train && fin_small) {
        lz.begin("init");
        launch_init(ia, h->init_w1_host, 0, h->sm_count, st);                                                         // partials -> partA
        ia.fin = make_fin(h->partA.as<double>(), np_init, h->init_bn[0], h->coef_init[0]); ia.part = h->partB.as<double>();
        launch_init(ia, h->init_w1_host, 1, h->sm_count, st);                                                         // finishes BN 0; partials -> partB
        ia.fin = make_fin(h->partB.as<double>(), np_init, h->init_bn[1], h->coef_init[1]); ia.part = nullptr;
        lz.end(2);
    } else if (train) {
        lz.begin("init"); launch_init(ia, h->init_w1_host, 0, h->sm_count, st); lz.end(1);
        lz.begin("bnfin"); finish_bn(h->partA.as<double>(), np_init, 32, h->init_bn[0], h->coef_init[0]); lz.end(h->world == 1 || h->px.ok ? 1 : 2);
        lz.begin("init"); launch_init(ia, h->init_w1_host, 1, h->sm_count, st); lz.end(1);
        lz.begin("bnfin"); finish_bn(h->partA.as<double>(), np_init, 32, h->init_bn[1], h->coef_init[1]); lz.end(h->world == 1 || h->px.ok ? 1 : 2);
    }
    lz.begin("init"); launch_init(ia, h->init_w1_host, 2, h->sm_count, st); lz.end(1);
    halo_exchange(h, h->mid[0]->as<float>(), nullptr, h->rflag(0), st, lz);

    // ---- message-passing layers ----------------------------------------------------------------
    const int n_layers = h->stop_layer >= 0 ? std::min(L, h->stop_layer + 1) : L;
    const int np_conv = conv_adj_num_parts(h->g.n_tiles, h->g.wn, h->sm_count, h->g.n_chunks);
    static const int64_t branch_max = getenv("TGNN_BRANCH_MAX") ? atoll(getenv("TGNN_BRANCH_MAX")) : BRANCH_MAX_NODES;
    const bool fork = !h->two_streams_off && h->world == 1 && !h->profiling && !h->role_dbg_on && h->g.n_own <= branch_max;
    if (fork && !h->side_stream) {
        TGNN_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
        TGNN_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        TGNN_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    for (int i = 0; i < n_layers; ++i) {
        const tgnn_handle::LayerP& P = h->lp[i];
        if (h->tables_streamed) { lz.begin("conv"); build_tables(h, st, i); lz.end(1); }
        const size_t tslot = h->tables_streamed ? 0 : (size_t)i * (h->g.n_types + 1);
        ConvArgs ca{};
        ca.xin = h->mid[i]->as<float>();
        ca.tabF = h->tab.as<float>() + tslot * TG_FRAG32;
        ca.n_types = h->g.n_types; ca.bias = P.conv_bias;
        ca.cptr = h->g.cptr.as<int>(); ca.ctype = h->g.ctype.as<int>(); ca.csrc = h->g.csrc.as<int>();
        ca.cdst = h->g.cdst.as<uint8_t>(); ca.inv_deg = h->mask_on ? h->inv_deg_m.as<float>() : h->g.inv_deg.as<float>();
        ca.mask = h->mask();
        ca.out = h->pre1.as<float>(); ca.part = train ? h->partA.as<double>() : nullptr;
        ca.n_own = n_own; ca.n_tiles = h->g.n_tiles; ca.wn = h->g.wn; ca.n_chunks = h->g.n_chunks;
        if (fork) {                                   // everything the two branches read is complete at this point of st
            TGNN_CUDA(cudaEventRecord(h->ev_fork, st));
            TGNN_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        }
        lz.begin("conv");
        if (h->use_z) {
            // fp16 kernel + its tf32 stand-by (exits at once unless a range flag is raised)
            ca.xh = h->xh.as<uint4>();
            ca.flag_x = h->rflag(i); ca.flag_w = h->wflag(i);
            launch_conv_z(ca, h->g, h->tabS.as<float>() + tslot * TG_FRAG32, h->tabT.as<uint32_t>() + tslot * TG_TIMG32, h->zflag(i),
                          h->conv_z32, h->err_dev, h->sm_count, st, h->role_dbg_on ? h->role_dbg.as<long long>() : nullptr);
            lz.end(h->conv_z32 ? 1 : 2, 1);
        } else if (h->use_s) {
            launch_conv_s(ca, h->g, h->tabS.as<float>() + tslot * TG_FRAG32, h->err_dev, h->sm_count, st);
            lz.end(1);
        } else if (h->use_t) {
            // tcgen05 edge-block kernel + its fp32 stand-by (exits at once unless a range flag is raised)
            ca.xh = h->xh.as<uint4>();
            ca.flag_x = h->rflag(i); ca.flag_w = h->wflag(i);
            launch_conv_t(ca, h->g, h->tabT.as<uint32_t>() + tslot * TG_TIMG32, h->tab32.as<float>() + tslot * F * F, h->err_dev,
                          h->sm_count, st, h->role_dbg_on ? h->role_dbg.as<long long>() : nullptr);
            lz.end(2, 1);
        } else if (h->use_h) {
            // fp16-split kernel; falls through to the 3xTF32 arithmetic itself when a range flag is raised
            ca.xh = h->xh.as<uint4>();
            ca.tabH = h->tabH.as<uint32_t>() + tslot * TG_HFRAG32;
            ca.flag_x = h->rflag(i); ca.flag_w = h->wflag(i);
            ca.dbg = h->role_dbg_on ? h->role_dbg.as<long long>() : nullptr;
            if (h->use_x) { ca.tabX = h->tabX.as<uint32_t>() + tslot * TG_HFRAG32; launch_conv_x(ca, h->sm_count, st); }
            else launch_conv_h(ca, h->sm_count, st);
            lz.end(1);
        } else {
            launch_conv_adj(ca, h->sm_count, st);
            lz.end(1);
        }

        GinArgs ga{};
        // pre2[0] = LeakyReLU(gin) before BatchNorm of this layer, pre2[1] = g2 = BN(pre2) written by k_combine
        ga.xin = i == 0 ? h->mid[0]->as<float>() : h->pre2[1].as<float>();
        ga.col_ptr = h->g.col_ptr.as<int>(); ga.col_src = h->g.col_src.as<int>();
        ga.wfrag = h->gin_wt[i]->as<float>();
        ga.eps = h->gin_eps[i]; ga.hmlp = h->gin_hmlp[i];
        ga.out = h->pre2[0].as<float>(); ga.part = train ? h->partB.as<double>() : nullptr; ga.n_own = n_own;
        ga.mask = h->mask();
        const bool gw = h->use_gw && ga.hmlp;        // layers whose GIN weights are outside the fp16 range stay on k_gin
        const int np_gin = gw ? gin_w_num_parts(h->g.gw_tiles, h->sm_count) : gin_num_parts(n_own, h->sm_count);
        lz.begin("gin");
        cudaStream_t st_gin = fork ? h->side_stream : st;
        if (gw) {
            ga.gw_meta = h->g.gw_meta.as<int>(); ga.gw_seg = h->g.gw_seg.as<int>(); ga.gw_loc = h->g.gw_loc.as<uint16_t>();
            ga.gw_tiles = h->g.gw_tiles; ga.err = h->err_dev;
            ga.dbg = h->role_dbg_on ? h->role_dbg.as<long long>() + 128 : nullptr;
            launch_gin_w(ga, h->sm_count, st_gin);
        } else launch_gin(ga, h->sm_count, st_gin);
        lz.end(1);
        if (fork) {
            TGNN_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
            TGNN_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
        }

        const int np_a = (h->use_s || h->use_z) ? h->g.s_tiles : (h->use_t ? conv_t_num_parts(h->g.t_tiles, h->sm_count) : np_conv);
        // small graphs: k_combine finishes the two BatchNorms in its prologue (one launch less per layer)
        const bool fin_in_combine = fin_small && np_a + np_gin <= 1024;
        CombineFin cf{};
        if (fin_in_combine) {
            cf.part[0] = h->partA.as<double>(); cf.n_part[0] = np_a; cf.part[1] = h->partB.as<double>(); cf.n_part[1] = np_gin;
            cf.count = count; cf.count_ptr = h->count_ptr();
            cf.gamma[0] = P.bn_a_w; cf.beta[0] = P.bn_a_b; cf.gamma[1] = P.bn_c_w; cf.beta[1] = P.bn_c_b;
            cf.coef_out[0] = h->C(h->coef_a[i]); cf.coef_out[1] = h->C(h->coef_c[i]);
        }
        if (train && !fin_in_combine) {
            lz.begin("bnfin");
            if (h->world == 1 || h->px.ok) {
                BnFinishArgs fa{};
                fa.part[0] = h->partA.as<double>(); fa.n_part[0] = np_a;
                fa.part[1] = h->partB.as<double>(); fa.n_part[1] = np_gin;
                fa.C = 32; fa.count = count;
                fa.gamma[0] = P.bn_a_w; fa.beta[0] = P.bn_a_b; fa.coef[0] = h->C(h->coef_a[i]);
                fa.gamma[1] = P.bn_c_w; fa.beta[1] = P.bn_c_b; fa.coef[1] = h->C(h->coef_c[i]);
                fa.sums = sums; fa.ticket = h->bn_ticket(); fa.count_ptr = h->count_ptr();
                if (h->world == 1) launch_bn_finish(fa, 2, st);
                else { launch_bn_finish_x(fa, 2, peer_ptrs(h), ++h->px.epoch_bn, st); h->collectives += 1; }
                lz.end(1);
            } else {
            launch_bn_reduce(h->partA.as<double>(), np_a, 32, sums, st);
            launch_bn_reduce(h->partB.as<double>(), np_gin, 32, sums + 64, st);
            allreduce_sums(h, sums, 128, st);
            launch_bn_coef(sums, count, P.bn_a_w, P.bn_a_b, h->C(h->coef_a[i]), 32, st);
            launch_bn_coef(sums + 64, count, P.bn_c_w, P.bn_c_b, h->C(h->coef_c[i]), 32, st);
            lz.end(4);
            }
        }
        lz.begin("combine");
        launch_combine(h->pre1.as<float>(), h->C(h->coef_a[i]), h->pre2[0].as<float>(), h->C(h->coef_c[i]),
                       i >= 2 ? h->mid[i - 2]->as<float>() : nullptr, h->mid[i + 1]->as<float>(),
                       h->need_xh() ? h->xh.as<uint4>() : nullptr, h->rflag(i + 1), h->pre2[1].as<float>(), n_own, st,
                       fin_in_combine ? &cf : nullptr, h->mask());
        lz.end(1);
        if (i + 1 < n_layers) halo_exchange(h, h->mid[i + 1]->as<float>(), h->pre2[1].as<float>(), h->rflag(i + 1), st, lz);
        h->last_layer_run = i;
    }

    // ---- final MLP -----------------------------------------------------------------------------
    if (h->stop_layer < 0) {
        int dims[5] = {F * (L + 1), 256, 128, 64, F};
        for (int k = 0; k < 4; ++k) {
            DenseArgs da{};
            da.slabs = k == 0 ? h->slab_ptrs.as<const float*>() : nullptr;
            da.a = k == 0 ? nullptr : h->fa[k - 1].as<float>();
            da.virtual_concat = k == 0;
            da.in_coef = k == 0 ? nullptr : h->C(h->coef_fin[k - 1]);
            da.wt = h->fin_wt[k]->as<float>(); da.bias = h->fin_bias[k];
            double* part_k = (k & 1) ? h->partB.as<double>() : h->partA.as<double>();
            double* part_prev = (k & 1) ? h->partA.as<double>() : h->partB.as<double>();
            const bool fin_k = fin_small && !h->dense_ffma;
            da.out = h->fa[k].as<float>(); da.part = train ? (fin_k ? part_k : h->partA.as<double>()) : nullptr;
            da.n = n_own; da.K = dims[k]; da.n_out = dims[k + 1]; da.mask = h->mask();
            if (fin_k && k > 0) da.fin = make_fin(part_prev, dense_row_blocks(n_own), h->fin_bn[k - 1], h->coef_fin[k - 1]);
            lz.begin("final");
            int nl = 1;
            if (h->dense_ffma) launch_dense(da, st);
            else {
                // fp16-split kernel + 3xTF32 stand-by (exits at once unless an input left the fp16 range)
                nl = launch_dense_tc(da, h->fin_whl[k]->as<float>(), h->fin_h16[k] ? h->fin_whh[k]->p : nullptr, h->dflag(k), h->err_dev,
                                     h->sm_count, st);
            }
            lz.end(nl, 1);
            if (train && !fin_k) {
                lz.begin("bnfin");
                finish_bn(h->partA.as<double>(), dense_row_blocks(n_own), dims[k + 1], h->fin_bn[k], h->coef_fin[k]);
                lz.end(h->world == 1 || h->px.ok ? 1 : 2);
            }
        }
        BnFin score_fin{};
        if (fin_small && !h->dense_ffma) score_fin = make_fin(h->partB.as<double>(), dense_row_blocks(n_own), h->fin_bn[3], h->coef_fin[3]);
        lz.begin("score");
        launch_score(h->fa[3].as<float>(), h->C(h->coef_fin[3]), h->score_w, h->fin_last_bias, scores,
                     n_own, st, h->mask(), score_fin.part ? &score_fin : nullptr);
        lz.end(1);
    }
    // Device-side errors are ALWAYS surfaced: synchronously here when that is cheap or asked for (small graphs: the
    // callers read the scores back at once anyway), otherwise at the next API call / tgnn_check_error (the error word
    // lives in mapped host memory, so reading it needs no CUDA call and no stall of the launch queue).
    if (!capturing && (h->profiling || h->check_errors || h->g.n_own <= SYNC_CHECK_MAX_NODES)) {
        TGNN_CUDA(cudaStreamSynchronize(st));
        check_device_error(h, "tgnn_forward");
    }
    if (h->profiling) {
        for (auto& e : h->prof) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e.a, e.b);
            h->prof_result[e.fam].first += ms;
        }
    }
}

constexpr int64_t REPLAY_MAX_NODES = 262144;     // above this a forward is milliseconds of kernels: launch overhead is noise

// tgnn_forward: replay the captured launch sequence when there is one, capture on the second use of a key, else eager.
void forward_entry(tgnn_handle* h, const float* x, float* scores, cudaStream_t st) {
    TGNN_CHECK(h->graph_set, "tgnn_forward: no graph set (call tgnn_set_graph first)");
    tgnn_handle::Replay& R = h->replay;
    const bool can = !R.disabled && h->world == 1 && !h->profiling && h->stop_layer < 0 && scores != nullptr &&
                     h->g.n_own <= REPLAY_MAX_NODES;
    if (!can) { forward_impl(h, x, scores, st); return; }
    check_device_error(h, "tgnn_forward (reported by an earlier forward)");
    if (h->params_dirty || h->tables_dirty) R.drop();                    // derived buffers may move: the captured pointers die
    pack_params(h, st);
    build_tables(h, st);
    if (h->tables_streamed) { forward_impl(h, x, scores, st); return; }   // per-layer table builds: not worth capturing
    const uint64_t key = (h->graph_gen << 20) ^ (h->param_gen << 3) ^ ((uint64_t)h->mask_on << 1) ^ (uint64_t)h->cfg.bn_mode;
    if (key != R.key) { R.drop(); R.key = key; }
    const size_t xb = (size_t)h->g.n_own * h->cfg.d_x * sizeof(float), sb = (size_t)h->g.n_own * sizeof(float);
    if (!R.exec) {
        if (++R.seen < 2) { forward_impl(h, x, scores, st); return; }
        // second forward on the same structures: capture it
        R.x_stage.reserve(xb); R.s_stage.reserve(sb);
        if (!R.cap_stream) TGNN_CUDA(cudaStreamCreateWithFlags(&R.cap_stream, cudaStreamNonBlocking));
        TGNN_CUDA(cudaStreamSynchronize(st));                             // table builds etc. issued on st are done
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(R.cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            try { forward_impl(h, R.x_stage.as<float>(), R.s_stage.as<float>(), R.cap_stream, true); }
            catch (const std::exception&) { ok = false; }
            if (cudaStreamEndCapture(R.cap_stream, &graph) != cudaSuccess) ok = false;
        }
        if (ok && graph) ok = cudaGraphInstantiate(&R.exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {                                                        // something in the sequence cannot be captured: stay eager
            cudaGetLastError();
            R.exec = nullptr; R.disabled = true;
            forward_impl(h, x, scores, st);
            return;
        }
        R.launches = h->launches;
    }
    TGNN_CUDA(cudaMemcpyAsync(R.x_stage.p, x, xb, cudaMemcpyDeviceToDevice, st));
    TGNN_CUDA(cudaGraphLaunch(R.exec, st));
    TGNN_CUDA(cudaMemcpyAsync(scores, R.s_stage.p, sb, cudaMemcpyDeviceToDevice, st));
    h->launches = R.launches; h->collectives = 0;
    if (h->train_mode_forward()) h->eval_coefs_valid = false;
    if (h->check_errors || h->g.n_own <= SYNC_CHECK_MAX_NODES) {
        TGNN_CUDA(cudaStreamSynchronize(st));
        check_device_error(h, "tgnn_forward");
    }
}

template <class Fn>
int guarded(tgnn_handle* h, Fn&& fn) {
    try {
        fn();
        return 0;
    } catch (const std::exception& e) {
        if (h) {
            h->err = e.what();
            // a kernel that aborted (trap) makes the next CUDA call fail with a generic message: say why
            if (h->err_host && *reinterpret_cast<volatile int*>(h->err_host) != 0) {
                h->err += std::string(" [device reported: ") + device_error_text(*reinterpret_cast<volatile int*>(h->err_host)) + "]";
                *reinterpret_cast<volatile int*>(h->err_host) = 0;
            }
        } else g_create_error = e.what();
        return 1;
    }
}

}  // namespace

extern "C" {

int tgnn_abi_version(void) { return TGNN_ABI_VERSION; }

int tgnn_create(const tgnn_cfg* cfg, tgnn_handle** out) {
    return guarded(nullptr, [&] {
        TGNN_CHECK(cfg && out, "tgnn_create: null argument");
        TGNN_CHECK(cfg->width == F, "tgnn_create: network_width must be 32");
        TGNN_CHECK(cfg->depth >= 1 && cfg->depth <= 64, "tgnn_create: network_depth must be in [1, 64]");
        TGNN_CHECK(cfg->d_x >= 1 && cfg->d_x <= 1024 && cfg->d_e >= 1 && cfg->d_e <= 4096, "tgnn_create: bad feature dims");
        TGNN_CHECK(cfg->bn_mode == TGNN_BN_TRAIN || cfg->bn_mode == TGNN_BN_EVAL, "tgnn_create: bad bn_mode");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        TGNN_CHECK(e == cudaSuccess && ndev > 0,
                   std::string("tgnn_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
        TGNN_CHECK(cfg->device >= 0 && cfg->device < ndev, "tgnn_create: bad device ordinal");
        DeviceGuard dg(cfg->device);
        cudaDeviceProp prop{};
        TGNN_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
        TGNN_CHECK(prop.major >= 10, "tgnn_create: this build targets sm_100a (B200); found sm_" +
                                         std::to_string(prop.major) + std::to_string(prop.minor));
        std::unique_ptr<tgnn_handle> h(new tgnn_handle());
        h->cfg = *cfg;
        h->sm_count = prop.multiProcessorCount;
        const char* csel = getenv("TGNN_CONV");
        h->conv_chunk_only = csel && std::string(csel) == "chunk";
        h->conv_s_only = csel && std::string(csel) == "s";
        h->conv_h_only = csel && (std::string(csel) == "h" || std::string(csel) == "x");
        h->conv_x_mode = !csel ? -1 : (std::string(csel) == "x" ? 1 : (std::string(csel) == "h" ? 0 : -1));
        h->conv_t_only = csel && std::string(csel) == "t";
        h->conv_z32 = csel && std::string(csel) == "z32";
        h->conv_z_only = csel && (std::string(csel) == "z" || h->conv_z32);
        const char* tsel = getenv("TGNN_TILE");
        if (tsel && (atoi(tsel) == WN_SMALL || atoi(tsel) == WN_BIG)) h->tile_rows_forced = atoi(tsel);
        h->hflags.reserve((size_t)(3 * cfg->depth + 3 + 4) * sizeof(int));
        TGNN_CUDA(cudaMemset(h->hflags.p, 0, (size_t)(3 * cfg->depth + 3 + 4) * sizeof(int)));
        h->dev_error.reserve(2 * sizeof(int));                         // [1] scratch flag of pack_params
        TGNN_CUDA(cudaMemset(h->dev_error.p, 0, 2 * sizeof(int)));
        TGNN_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->err_host), 64, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(h->err_host, 0, 64);
        TGNN_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->err_dev), h->err_host, 0));
        declare_params(h.get());
        *out = h.release();
    });
}

int tgnn_destroy(tgnn_handle* h) {
    if (!h) return 0;
    {
        DeviceGuard dg(h->cfg.device);
        for (auto& e : h->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        h->replay.drop();
        if (h->replay.cap_stream) cudaStreamDestroy(h->replay.cap_stream);
        if (h->side_stream) cudaStreamDestroy(h->side_stream);
        if (h->ev_fork) cudaEventDestroy(h->ev_fork);
        if (h->ev_join) cudaEventDestroy(h->ev_join);
        peer_close(h);
        if (h->comm && nccl().CommDestroy) nccl().CommDestroy(h->comm);
        if (h->err_host) cudaFreeHost(h->err_host);
        delete h;
    }
    return 0;
}

int tgnn_set_param(tgnn_handle* h, const char* ref_key, const void* data, const int64_t* shape, int32_t ndim) {
    return guarded(h, [&] {
        TGNN_CHECK(h && ref_key, "tgnn_set_param: null argument");
        std::string key(ref_key);
        if (ends_with(key, ".num_batches_tracked")) return;            // int64 counter, irrelevant to outputs
        key = canonical_key(key);
        auto it = h->params.find(key);
        TGNN_CHECK(it != h->params.end(), "tgnn_set_param: unexpected key " + std::string(ref_key));
        Param& p = it->second;
        TGNN_CHECK(data != nullptr, "tgnn_set_param: null data for " + key);
        bool ok = (int)p.shape.size() == ndim;
        for (int i = 0; ok && i < ndim; ++i) ok = p.shape[i] == shape[i];
        if (!ok) {
            std::string want, got;
            for (auto d : p.shape) want += std::to_string(d) + ",";
            for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
            throw Error("tgnn_set_param: size mismatch for " + key + ": expected [" + want + "] got [" + got + "]");
        }
        DeviceGuard dg(h->cfg.device);
        size_t bytes = p.numel() * sizeof(float);
        p.buf.reserve(bytes);
        TGNN_CUDA(cudaMemcpy(p.buf.p, data, bytes, cudaMemcpyDefault));
        p.set = true;
        h->params_dirty = true;
        h->param_gen++;
    });
}

int tgnn_missing_params(tgnn_handle* h, char* buf, int32_t buflen) {
    if (!h) return -1;
    int missing = 0;
    for (auto& k : h->key_order)
        if (!h->params[k].set) {
            if (missing == 0 && buf && buflen > 0) { strncpy(buf, k.c_str(), buflen - 1); buf[buflen - 1] = 0; }
            ++missing;
        }
    return missing;
}

int tgnn_set_bn_mode(tgnn_handle* h, int32_t bn_mode) {
    return guarded(h, [&] {
        TGNN_CHECK(bn_mode == TGNN_BN_TRAIN || bn_mode == TGNN_BN_EVAL, "tgnn_set_bn_mode: bad mode");
        h->cfg.bn_mode = bn_mode;
    });
}

int tgnn_set_graph(tgnn_handle* h, int64_t n_nodes, int64_t e_adj, const int64_t* adj_src, const int64_t* adj_dst,
                   const float* adj_feat, int64_t e_col, const int64_t* col_src, const int64_t* col_dst, void* stream) {
    return guarded(h, [&] {
        TGNN_CHECK(h, "tgnn_set_graph: null handle");
        TGNN_CHECK(h->world == 1, "tgnn_set_graph: handle is sharded, use tgnn_set_graph_shard");
        check_device_error(h, "tgnn_set_graph (reported by an earlier forward)");
        DeviceGuard dg(h->cfg.device);
        cudaStream_t st = (cudaStream_t)stream;
        h->graph_set = false;
        build_graph(h->g, h->scratch, h->cfg.d_e, n_nodes, n_nodes, e_adj, adj_src, adj_dst, adj_feat, e_col, col_src,
                    col_dst, want_s_mode(h), tile_rows_for(h, n_nodes), want_t_rows(h, n_nodes), st);
        h->g.n_global = n_nodes; h->g.halo_slot = 0; h->g.n_send = 0;
        choose_conv_kernel(h, st);
        choose_gin_kernel(h, st);
        alloc_workspace(h);
        {
            const bool was_valid = !h->tables_dirty && !h->params_dirty;
            const uint64_t sig = ((uint64_t)h->g.n_types << 8) | ((uint64_t)h->use_h << 0) | ((uint64_t)h->use_x << 1) | ((uint64_t)h->use_s << 2) |
                                 ((uint64_t)h->use_t << 3) | ((uint64_t)h->use_z << 4) | ((uint64_t)h->tables_streamed << 5);
            bool same = false;
            if (h->g.n_types > 0 && h->g.n_types <= 1024) {
                std::vector<float> rows((size_t)h->g.n_types * h->cfg.d_e);
                TGNN_CUDA(cudaMemcpyAsync(rows.data(), h->g.type_rows.p, rows.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
                TGNN_CUDA(cudaStreamSynchronize(st));
                same = was_valid && sig == h->tables_sig && rows.size() == h->type_rows_host.size() &&
                       memcmp(rows.data(), h->type_rows_host.data(), rows.size() * sizeof(float)) == 0;
                h->type_rows_host.swap(rows);
            } else h->type_rows_host.clear();
            h->tables_sig = sig;
            h->tables_dirty = !same;
        }
        h->graph_gen++;
        h->mask_on = false;
        h->graph_set = true;
    });
}

int tgnn_forward(tgnn_handle* h, const float* x, float* scores_out, void* stream) {
    return guarded(h, [&] {
        TGNN_CHECK(h && x && (scores_out || h->stop_layer >= 0), "tgnn_forward: null argument");
        DeviceGuard dg(h->cfg.device);
        forward_entry(h, x, scores_out, (cudaStream_t)stream);
    });
}

int tgnn_nccl_unique_id(void* out128) {
    return guarded(nullptr, [&] {
        TGNN_CHECK(nccl().GetUniqueId, "NCCL is not available in this process (libnccl.so.2 not found)");
        ncclUniqueId id;
        nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out128, &id, 128);
    });
}

int tgnn_shard_init(tgnn_handle* h, const void* unique_id128, int32_t rank, int32_t world) {
    return guarded(h, [&] {
        TGNN_CHECK(h && unique_id128 && world >= 1 && rank >= 0 && rank < world, "tgnn_shard_init: bad arguments");
        h->rank = rank; h->world = world;
        if (world == 1) return;
        TGNN_CHECK(nccl().CommInitRank, "NCCL is not available in this process (libnccl.so.2 not found)");
        DeviceGuard dg(h->cfg.device);
        ncclUniqueId id;
        memcpy(&id, unique_id128, 128);
        nccl_check(nccl().CommInitRank(&h->comm, world, id, rank), "ncclCommInitRank");
    });
}

int tgnn_set_graph_shard(tgnn_handle* h, int64_t n_own, int64_t n_global, int64_t halo_slot, int64_t n_send,
                         const int64_t* send_rows, int64_t e_adj, const int64_t* adj_src, const int64_t* adj_dst,
                         const float* adj_feat, int64_t e_col, const int64_t* col_src, const int64_t* col_dst, void* stream) {
    return guarded(h, [&] {
        TGNN_CHECK(h, "tgnn_set_graph_shard: null handle");
        TGNN_CHECK(halo_slot >= 0 && n_send >= 0 && n_send <= halo_slot && n_global >= n_own, "tgnn_set_graph_shard: bad sizes");
        check_device_error(h, "tgnn_set_graph_shard (reported by an earlier forward)");
        DeviceGuard dg(h->cfg.device);
        cudaStream_t st = (cudaStream_t)stream;
        h->graph_set = false;
        const int64_t n_rows = n_own + (h->world > 1 ? (int64_t)h->world * halo_slot : 0);
        build_graph(h->g, h->scratch, h->cfg.d_e, n_own, n_rows, e_adj, adj_src, adj_dst, adj_feat, e_col, col_src, col_dst,
                    want_s_mode(h), tile_rows_for(h, n_own), want_t_rows(h, n_own), st);
        h->g.n_global = n_global; h->g.halo_slot = halo_slot; h->g.n_send = n_send;
        choose_conv_kernel(h, st);
        choose_gin_kernel(h, st);
        if (n_send > 0) {
            std::vector<int64_t> rows64(n_send);
            TGNN_CUDA(cudaMemcpyAsync(rows64.data(), send_rows, n_send * sizeof(int64_t), cudaMemcpyDefault, st));
            TGNN_CUDA(cudaStreamSynchronize(st));
            std::vector<int> rows32(n_send);
            for (int64_t i = 0; i < n_send; ++i) {
                TGNN_CHECK(rows64[i] >= 0 && rows64[i] < n_own, "tgnn_set_graph_shard: send row out of range");
                rows32[i] = (int)rows64[i];
            }
            h->g.send_rows.reserve(n_send * sizeof(int));
            TGNN_CUDA(cudaMemcpyAsync(h->g.send_rows.p, rows32.data(), n_send * sizeof(int), cudaMemcpyHostToDevice, st));
            TGNN_CUDA(cudaStreamSynchronize(st));
        }
        // which mirrored rows do the local edges read, and which peers own them: the unpack waits for those peers only
        h->g.has_send_mask = false; h->g.need_from = 0xffffffffu;
        if (h->world > 1 && halo_slot > 0) {
            const size_t nm = (size_t)h->world * (size_t)halo_slot;
            h->g.halo_used.reserve(nm);
            TGNN_CUDA(cudaMemsetAsync(h->g.halo_used.p, 0, nm, st));
            launch_mark_halo(adj_src, e_adj, n_own, n_rows, h->g.halo_used.as<uint8_t>(), st);
            launch_mark_halo(col_src, e_col, n_own, n_rows, h->g.halo_used.as<uint8_t>(), st);
            std::vector<uint8_t> used(nm);
            TGNN_CUDA(cudaMemcpyAsync(used.data(), h->g.halo_used.p, nm, cudaMemcpyDeviceToHost, st));
            TGNN_CUDA(cudaStreamSynchronize(st));
            unsigned nf = 0;
            for (int q = 0; q < h->world; ++q)
                for (int64_t i = 0; i < halo_slot; ++i)
                    if (used[(size_t)q * halo_slot + i]) { nf |= 1u << q; break; }
            h->g.need_from = nf;
        }
        alloc_workspace(h);
        peer_setup(h, st);
        h->tables_dirty = true;
        h->graph_set = true;
    });
}

int tgnn_set_halo_peers(tgnn_handle* h, const uint8_t* send_mask, int64_t n_send, void* stream) {
    return guarded(h, [&] {
        TGNN_CHECK(h && h->graph_set && h->world > 1, "tgnn_set_halo_peers: no sharded graph set");
        TGNN_CHECK(n_send == h->g.n_send, "tgnn_set_halo_peers: one mask byte per send row");
        TGNN_CHECK(h->world <= 8, "tgnn_set_halo_peers: the mask has one bit per rank (world <= 8)");
        DeviceGuard dg(h->cfg.device);
        cudaStream_t st = (cudaStream_t)stream;
        if (n_send > 0) {
            TGNN_CHECK(send_mask, "tgnn_set_halo_peers: null mask");
            h->g.send_mask.reserve((size_t)n_send);
            TGNN_CUDA(cudaMemcpyAsync(h->g.send_mask.p, send_mask, (size_t)n_send, cudaMemcpyDefault, st));
            TGNN_CUDA(cudaStreamSynchronize(st));
        }
        h->g.has_send_mask = n_send > 0;
    });
}

int tgnn_get_info(tgnn_handle* h, tgnn_info* out) {
    return guarded(h, [&] {
        TGNN_CHECK(h && out, "tgnn_get_info: null argument");
        out->n_own = h->g.n_own; out->n_rows = h->g.n_rows; out->n_global = h->g.n_global;
        out->e_adj = h->g.e_adj; out->e_col = h->g.e_col; out->n_edge_types = h->g.n_types;
        out->adj_slots = (int64_t)h->g.n_chunks * CH;
        out->launches_per_forward = h->launches;
        out->workspace_bytes = (int64_t)h->workspace_bytes;
        out->collectives_per_forward = h->collectives;
        out->conv_kernel = h->use_z ? 4 : (h->use_s ? 1 : (h->use_t ? 3 : (h->use_x ? 5 : (h->use_h ? 2 : 0))));
        out->tile_rows = h->g.wn;
        out->peer_exchange = h->px.ok ? 1 : 0;
        out->t_rows = h->g.has_t ? h->g.t_rows : 0;
        out->t_blocks = h->g.has_t ? h->g.t_blocks : 0;
        out->gin_kernel = h->use_gw ? 1 : 0;
        out->gin_window_tiles = h->g.has_gw ? h->g.gw_tiles - h->g.gw_direct : 0;
        out->gin_direct_tiles = h->g.has_gw ? h->g.gw_direct : 0;
        out->range_fallback_layers = 0;
        if (h->need_xh() && h->graph_set) {
            DeviceGuard dg(h->cfg.device);
            const int L = h->cfg.depth;
            std::vector<int> f(3 * L + 3);
            TGNN_CUDA(cudaDeviceSynchronize());
            TGNN_CUDA(cudaMemcpy(f.data(), h->hflags.p, f.size() * sizeof(int), cudaMemcpyDeviceToHost));
            for (int i = 0; i < L; ++i) out->range_fallback_layers += (f[i] | f[L + 1 + i] | (h->use_z ? f[2 * L + 3 + i] : 0)) ? 1 : 0;
        }
    });
}

int tgnn_set_node_mask(tgnn_handle* h, const uint8_t* keep, int64_t* counts3, void* stream) {
    return guarded(h, [&] {
        TGNN_CHECK(h && h->graph_set, "tgnn_set_node_mask: no graph set");
        TGNN_CHECK(h->world == 1, "tgnn_set_node_mask: not supported on a sharded handle");
        DeviceGuard dg(h->cfg.device);
        cudaStream_t st = (cudaStream_t)stream;
        if (!keep) {
            h->mask_on = false;
            if (counts3) { counts3[0] = h->g.n_own; counts3[1] = h->g.e_adj; counts3[2] = h->g.e_col; }
            return;
        }
        const size_t n = (size_t)h->g.n_own;
        const void *b0 = h->mask_buf.p, *b1 = h->inv_deg_m.p, *b2 = h->mask_cnt.p;
        h->mask_buf.reserve(n); h->inv_deg_m.reserve(n * sizeof(float)); h->mask_cnt.reserve(32);
        if (b0 != h->mask_buf.p || b1 != h->inv_deg_m.p || b2 != h->mask_cnt.p) h->replay.drop();     // captured pointers moved
        TGNN_CUDA(cudaMemcpyAsync(h->mask_buf.p, keep, n, cudaMemcpyDefault, st));
        launch_node_mask(h->g, h->mask_buf.as<uint8_t>(), h->inv_deg_m.as<float>(), h->mask_cnt.as<int>(),
                         reinterpret_cast<double*>(h->mask_cnt.as<char>() + 16), st);
        h->mask_on = true;
        if (counts3) {
            int c[3];
            TGNN_CUDA(cudaMemcpyAsync(c, h->mask_cnt.p, sizeof(c), cudaMemcpyDeviceToHost, st));
            TGNN_CUDA(cudaStreamSynchronize(st));
            counts3[0] = c[0]; counts3[1] = c[1]; counts3[2] = c[2];
        }
    });
}

int tgnn_check_error(tgnn_handle* h, void* stream, int32_t synchronize) {
    return guarded(h, [&] {
        TGNN_CHECK(h, "tgnn_check_error: null handle");
        if (synchronize) {
            DeviceGuard dg(h->cfg.device);
            TGNN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        }
        check_device_error(h, "tgnn_check_error");
    });
}

int tgnn_debug_set_stop_layer(tgnn_handle* h, int32_t layer) {
    return guarded(h, [&] { h->stop_layer = layer; });
}

int tgnn_debug_read(tgnn_handle* h, const char* name, float* out, void* stream) {
    return guarded(h, [&] {
        TGNN_CHECK(h && name && out && h->graph_set, "tgnn_debug_read: bad arguments");
        DeviceGuard dg(h->cfg.device);
        cudaStream_t st = (cudaStream_t)stream;
        std::string n(name);
        const size_t bytes = (size_t)h->g.n_own * F * sizeof(float);
        const int i = h->last_layer_run;
        if (n.rfind("mid_", 0) == 0) {
            int k = std::stoi(n.substr(4));
            TGNN_CHECK(k >= 0 && k <= h->cfg.depth, "tgnn_debug_read: bad slab index");
            TGNN_CUDA(cudaMemcpyAsync(out, h->mid[k]->p, bytes, cudaMemcpyDeviceToDevice, st));
        } else if (n == "pre1") {
            TGNN_CUDA(cudaMemcpyAsync(out, h->pre1.p, bytes, cudaMemcpyDeviceToDevice, st));
        } else if (n == "pre2") {
            TGNN_CHECK(i >= 0, "tgnn_debug_read: no layer has run");
            TGNN_CUDA(cudaMemcpyAsync(out, h->pre2[0].p, bytes, cudaMemcpyDeviceToDevice, st));
        } else if (n == "g1" || n == "g2") {
            // BN(pre) of the last layer that ran: combine with the other factor's BN replaced by identity
            TGNN_CHECK(i >= 0, "tgnn_debug_read: no layer has run");
            DevBuf ident;
            ident.reserve(128 * sizeof(float));
            std::vector<float> idc(128, 0.f);
            for (int c = 0; c < 32; ++c) idc[64 + c] = 0.f, idc[96 + c] = 1.f;     // scale 0, beta 1 -> factor 1
            TGNN_CUDA(cudaMemcpyAsync(ident.p, idc.data(), 128 * sizeof(float), cudaMemcpyHostToDevice, st));
            if (n == "g1")
                launch_combine(h->pre1.as<float>(), h->C(h->coef_a[i]), h->pre2[0].as<float>(), ident.as<float>(), nullptr, out,
                               nullptr, nullptr, nullptr, h->g.n_own, st);
            else
                launch_combine(h->pre1.as<float>(), ident.as<float>(), h->pre2[0].as<float>(), h->C(h->coef_c[i]), nullptr, out,
                               nullptr, nullptr, nullptr, h->g.n_own, st);
            TGNN_CUDA(cudaStreamSynchronize(st));
        } else {
            throw Error("tgnn_debug_read: unknown tensor " + n);
        }
    });
}

int tgnn_debug_graph(tgnn_handle* h, int32_t* cptr, int32_t* ctype, int32_t* csrc, uint8_t* cdst, float* inv_deg,
                     int32_t* col_ptr, int32_t* col_src, float* type_rows, void* stream) {
    // Copies the built graph structures to caller-provided DEVICE or HOST buffers (sizes from tgnn_get_info and the
    // struct Graph comments); null pointers are skipped.  Test-only.
    return guarded(h, [&] {
        TGNN_CHECK(h && h->graph_set, "tgnn_debug_graph: no graph");
        DeviceGuard dg(h->cfg.device);
        cudaStream_t st = (cudaStream_t)stream;
        auto cp = [&](void* dst, const DevBuf& b, size_t bytes) {
            if (dst && bytes) TGNN_CUDA(cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyDefault, st));
        };
        cp(cptr, h->g.cptr, (size_t)(h->g.n_tiles + 1) * 4);
        cp(ctype, h->g.ctype, (size_t)h->g.n_chunks * 4);
        cp(csrc, h->g.csrc, (size_t)h->g.n_chunks * CH * 4);
        cp(cdst, h->g.cdst, (size_t)h->g.n_chunks * CH);
        cp(inv_deg, h->g.inv_deg, (size_t)h->g.n_own * 4);
        cp(col_ptr, h->g.col_ptr, (size_t)(h->g.n_own + 1) * 4);
        cp(col_src, h->g.col_src, (size_t)h->g.e_col * 4);
        cp(type_rows, h->g.type_rows, (size_t)h->g.n_types * h->cfg.d_e * 4);
        TGNN_CUDA(cudaStreamSynchronize(st));
    });
}

int tgnn_debug_graph_t(tgnn_handle* h, int32_t* bptr, int32_t* btype, int32_t* tsrc, uint16_t* tdst, void* stream) {
    return guarded(h, [&] {
        TGNN_CHECK(h && h->graph_set && h->g.has_t, "tgnn_debug_graph_t: no edge-block format on this graph");
        DeviceGuard dg(h->cfg.device);
        cudaStream_t st = (cudaStream_t)stream;
        auto cp = [&](void* dst, const DevBuf& b, size_t bytes) {
            if (dst && bytes) TGNN_CUDA(cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyDefault, st));
        };
        cp(bptr, h->g.t_bptr, (size_t)(h->g.t_tiles + 1) * 4);
        cp(btype, h->g.t_btype, (size_t)h->g.t_blocks * 4);
        cp(tsrc, h->g.t_src, (size_t)h->g.t_blocks * 128 * 4);
        cp(tdst, h->g.t_dst, (size_t)h->g.t_blocks * 128 * 2);
        TGNN_CUDA(cudaStreamSynchronize(st));
    });
}

int tgnn_debug_role_cycles(tgnn_handle* h, int64_t* out256) {
    // TGNN_ROLE_DBG=1: per-warp {cycles, wait 0, wait 1, wait 2} of CTA 0 in the last launch of k_conv_t ([0,128)) and of
    // k_gin_w ([128,256)); synchronises the device
    return guarded(h, [&] {
        TGNN_CHECK(h && out256 && h->role_dbg.p, "tgnn_debug_role_cycles: role timing is off (set TGNN_ROLE_DBG=1 before tgnn_create)");
        DeviceGuard dg(h->cfg.device);
        TGNN_CUDA(cudaDeviceSynchronize());
        TGNN_CUDA(cudaMemcpy(out256, h->role_dbg.p, 256 * sizeof(long long), cudaMemcpyDeviceToHost));
    });
}

int tgnn_set_profiling(tgnn_handle* h, int32_t enabled) {
    return guarded(h, [&] { h->profiling = enabled != 0; });
}

int tgnn_get_profile(tgnn_handle* h, const char* name, float* ms_out, int32_t* launches_out) {
    return guarded(h, [&] {
        TGNN_CHECK(h && name, "tgnn_get_profile: null argument");
        auto it = h->prof_result.find(name);
        float ms = 0.f; int n = 0;
        if (it != h->prof_result.end()) { ms = it->second.first; n = it->second.second; }
        if (ms_out) *ms_out = ms;
        if (launches_out) *launches_out = n;
    });
}

const char* tgnn_last_error(tgnn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

}  // extern "C"
