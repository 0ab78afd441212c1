/*
 * tgnn.h -- C ABI of the B200-native TilinGNN scoring path (libtgnn.so).
 *
 * The reference (xuhaocuhk/TilinGNN) has no FFI / plugin registry: its boundary
 * for this path is a Python nn.Module call.  This header is the C-ABI a
 * maintainer binds in its place (ctypes stub in INTEGRATION.md); every entry
 * point cites the reference interface it replaces (paths relative to the
 * reference root).  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from tgnn_last_error(h) (or tgnn_last_error(NULL) for failures
 *     of tgnn_create itself).  The Python wrapper raises RuntimeError with it,
 *     which preserves the reference's contract that a failed forward surfaces
 *     as a Python exception (graph_networks/network_utils.py:10-19).
 *   - the caller owns every buffer passed in; the handle owns parameters,
 *     graph structure, workspace.  Inputs are never written.
 *   - all device work is ordered on the `stream` argument (a cudaStream_t
 *     passed as void*; NULL = the legacy default stream).
 *   - one handle is NOT thread-safe; distinct handles are independent.
 *   - there is no CPU fallback: without a CUDA device tgnn_create fails.
 */
#ifndef TGNN_H_
#define TGNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGNN_ABI_VERSION 5

#define TGNN_BN_TRAIN 0   /* batch statistics over the rows of THIS call -- the reference's
                             behaviour: solver/ml_solver/ml_solver.py:129-131 ends in network.train() */
#define TGNN_BN_EVAL  1   /* running statistics from the checkpoint (nn.Module.eval())               */

typedef struct tgnn_handle tgnn_handle;

/* Constructor arguments of TilinGNN.__init__ (graph_networks/networks/TilinGNN.py:14-20).
 * width must be 32 (inputs/config.py:38; every shipped checkpoint). */
typedef struct tgnn_cfg {
    int32_t d_x;        /* node_features_dim   = environment.tile_count + 1            */
    int32_t d_e;        /* adj_edge_features_dim                                       */
    int32_t width;      /* network_width (32)                                          */
    int32_t depth;      /* network_depth (20 shipped; 6 in the benchmark configs)      */
    int32_t bn_mode;    /* TGNN_BN_TRAIN | TGNN_BN_EVAL                                */
    int32_t device;     /* CUDA device ordinal                                         */
} tgnn_cfg;

int tgnn_abi_version(void);

/* TilinGNN(...)  -- graph_networks/networks/TilinGNN.py:14-48 */
int tgnn_create(const tgnn_cfg* cfg, tgnn_handle** out);
int tgnn_destroy(tgnn_handle* h);

/* network.load_state_dict(torch.load(path)) -- solver/ml_solver/ml_solver.py:129-130.
 * One call per reference state_dict key (fp32 data, host or device pointer; shape checked).
 * The aliased "...nnConv.nn.mlp.*" keys (graph_networks/layers/edge_conv.py:17-18 registers one
 * MLP under two names) address the same tensor as "...mlp.mlp.*" (the last write wins);
 * "...num_batches_tracked" keys are accepted and ignored (int64). */
int tgnn_set_param(tgnn_handle* h, const char* ref_key, const void* data,
                   const int64_t* shape, int32_t ndim);
/* Number of reference keys still unset (0 = ready); writes the first missing key to buf. */
int tgnn_missing_params(tgnn_handle* h, char* buf, int32_t buflen);
/* nn.Module.train() / .eval() */
int tgnn_set_bn_mode(tgnn_handle* h, int32_t bn_mode);

/* The graph arguments of TilinGNN.forward (graph_networks/networks/TilinGNN.py:51):
 * adj_e_index = [adj_src; adj_dst] and col_e_idx = [col_src; col_dst] in PyG
 * source_to_target order, int64, as util/data_util.py:110-117 produces them;
 * adj_feat = adj_e_features [E_a, d_e] fp32 row-major.  Device pointers.
 * Builds the device-side structures (edge-type ids, typed adjacency tiles, collision CSR,
 * per-layer edge-weight tables).  Collision self loops are dropped (PyG GINConv).
 * Edge types = distinct adj_feat rows (bitwise, -0 == +0); any number up to 2^22 is accepted -- with
 * many types (continuous features) the weight tables are built one layer at a time. */
int tgnn_set_graph(tgnn_handle* h, int64_t n_nodes,
                   int64_t e_adj, const int64_t* adj_src, const int64_t* adj_dst, const float* adj_feat,
                   int64_t e_col, const int64_t* col_src, const int64_t* col_dst,
                   void* stream);

/* TilinGNN.forward(x, ...) -> scores  (graph_networks/networks/TilinGNN.py:51-78).
 * x: [n_nodes, d_x] fp32 device; scores_out: [n_nodes] fp32 device (the [N,1] column). */
int tgnn_forward(tgnn_handle* h, const float* x, float* scores_out, void* stream);

/* BrickLayout.compute_sub_layout (tiling/brick_layout.py:248-286) WITHOUT re-indexing: the sub-layout of the resident
 * graph induced by the nodes with keep[i] != 0 (uint8 [n_nodes], host or device pointer; NULL = all nodes again).
 * The next forwards score exactly that sub-graph on the structures already in HBM: masked rows are zero in every
 * tensor a neighbour gathers (so sums over all neighbours are sums over the kept ones), mean aggregation divides by
 * the kept in-degree, BatchNorm statistics run over the kept nodes only.  scores_out[i] of a masked node is 0.
 * A greedy round (util/algorithms.py:27-31) is then n_nodes bytes of upload + one forward, no rebuild.
 * counts3 (host, optional; synchronises): {kept nodes, adjacency edges, collision edges with both endpoints kept} --
 * what ML_Solver.predict's early-out (solver/ml_solver/ml_solver.py:31-32) needs.  Single-GPU handles only; a new
 * tgnn_set_graph clears the mask. */
int tgnn_set_node_mask(tgnn_handle* h, const uint8_t* keep, int64_t* counts3, void* stream);

/* Device-side failures (a tcgen05 pipeline timeout, a peer-exchange wait that timed out) are recorded by
 * the kernels in a word of mapped host memory.  They are always surfaced: by tgnn_forward itself when the
 * graph is small (it then synchronises `stream`, the callers read the scores back at once anyway), otherwise
 * at the start of the next API call on the handle -- or here.  synchronize != 0: wait for `stream` first (what
 * a caller does before trusting scores of a large graph); 0: host read only.  Non-zero return = an error was
 * pending (tgnn_last_error has the text; the word is cleared).  Keeps the reference's contract that a failed
 * forward surfaces as a Python exception (graph_networks/network_utils.py:10-19). */
int tgnn_check_error(tgnn_handle* h, void* stream, int32_t synchronize);

/* ---- multi-GPU: node-range shards (new work; the reference is single-device) ------------
 * Per layer one exchange of boundary rows and one of BatchNorm sums.  By default both are peer-memory
 * exchanges: tgnn_set_graph_shard maps every peer's exchange buffer with CUDA IPC (handles travel through
 * a 64-byte NCCL all-gather) and the producing kernels store straight into the peers' buffers over NVLink;
 * if the mapping fails on any rank, or with TGNN_P2P=0, ncclAllGather / ncclAllReduce are used instead. */
/* 128-byte NCCL unique id; rank 0 creates it, the host side broadcasts it. */
int tgnn_nccl_unique_id(void* out128);
/* Join the communicator.  Must be called before tgnn_set_graph_shard. */
int tgnn_shard_init(tgnn_handle* h, const void* unique_id128, int32_t rank, int32_t world);
/* Sharded graph in LOCAL row numbering: rows [0, n_own) are this rank's nodes, rows
 * [n_own, n_own + world*halo_slot) mirror the boundary rows every rank publishes (slot r holds
 * rank r's send list, padded to halo_slot rows).  dst in [0, n_own), src in [0, n_rows).
 * send_rows[n_send] (n_send <= halo_slot): this rank's own rows that peers read.
 * n_global = total node count over all ranks (BatchNorm statistics are global). */
int tgnn_set_graph_shard(tgnn_handle* h, int64_t n_own, int64_t n_global, int64_t halo_slot,
                         int64_t n_send, const int64_t* send_rows,
                         int64_t e_adj, const int64_t* adj_src, const int64_t* adj_dst, const float* adj_feat,
                         int64_t e_col, const int64_t* col_src, const int64_t* col_dst,
                         void* stream);
/* Optional, after tgnn_set_graph_shard: send_mask[n_send] (host or device), bit q of byte r = rank q reads send row r.
 * The peer-memory exchange then stores a boundary row only into the buffers of the ranks that read it (node-range shards of
 * a spatially ordered graph have two neighbours, not world - 1); which mirrored rows THIS rank reads, and therefore whose
 * flags it waits for, the library derives from the edge arrays itself.  Without this call every row goes to every peer.
 * (The NCCL fallback all-gathers everything either way.)  New work: the reference is single-device. */
int tgnn_set_halo_peers(tgnn_handle* h, const uint8_t* send_mask, int64_t n_send, void* stream);

/* ---- introspection (tests, bench, profiling) --------------------------------------------- */
typedef struct tgnn_info {
    int64_t n_own, n_rows, n_global;
    int64_t e_adj, e_col;          /* e_col after self-loop removal                          */
    int64_t n_edge_types;          /* K distinct adjacency feature rows                      */
    int64_t adj_slots;             /* padded slots of the typed adjacency tiles (>= e_adj)   */
    int64_t launches_per_forward;  /* kernels of this library launched by one tgnn_forward   */
    int64_t workspace_bytes;
    int64_t collectives_per_forward;
    int64_t conv_kernel;           /* adjacency kernel chosen for this graph: 0 = 3xTF32 edge-chunk (mma.sync),
                                      1 = tcgen05 S formulation, 2 = fp16-split edge-chunk (mma.sync.f16),
                                      3 = tcgen05 edge-block kernel (128-edge blocks, TMEM accumulators),
                                      4 = windowed tcgen05 kernel (A operand in tensor memory, rows staged per tile),
                                      5 = fp16-split edge-chunk kernel with transposed MMA roles (weights = A operand)  */
    int64_t tile_rows;             /* destination rows per warp tile of the typed adjacency format (64 or 128)  */
    int64_t peer_exchange;         /* sharded mode: 1 = boundary rows and BatchNorm sums travel as direct NVLink stores into
                                      the peers' CUDA-IPC-mapped buffers (flags, no NCCL call); 0 = NCCL collectives     */
    int64_t t_rows, t_blocks;      /* edge-block format of kernel 3 (0 when not built): destination rows per super-tile,
                                      blocks of 128 same-type slots (incl. t_rows / 128 root blocks per super-tile)         */
    int64_t gin_kernel;            /* collision kernel chosen for this graph: 0 = per-lane global gathers (k_gin),
                                      1 = neighbour rows staged in shared-memory windows by TMA bulk copies (k_gin_w)  */
    int64_t gin_window_tiles;      /* 64-row tiles that got a window / that are gathered from global ("direct")        */
    int64_t gin_direct_tiles;
    int64_t range_fallback_layers; /* layers of the LAST forward that kernel 2 handed to kernel 0 because an
                                      activation or root weight was outside the fp16 range (synchronises)      */
} tgnn_info;
int tgnn_get_info(tgnn_handle* h, tgnn_info* out);

/* Test hooks.  tgnn_debug_set_stop_layer(h, i >= 0): the next forwards stop after message-passing
 * layer i (no final MLP, scores untouched); -1 restores the full forward.
 * tgnn_debug_read copies an internal tensor of the LAST forward to `out` (device, fp32, [n_own,32]):
 *   "mid_<k>"  middle_features[k] of TilinGNN.forward (k = 0 is the init MLP output, k = i+1 layer i's b1)
 *   "pre1" / "pre2"  LeakyReLU(conv) of the last layer that ran, before BatchNorm
 *   "g1" / "g2"      GraphConv / CollConv output (after BatchNorm) of the last layer that ran
 * tgnn_debug_graph copies the built graph structures out (sizes: tgnn_get_info; null = skip):
 *   cptr[n_tiles+1] ctype[n_chunks] csrc[adj_slots] cdst[adj_slots] inv_deg[n_own]
 *   col_ptr[n_own+1] col_src[e_col] type_rows[n_edge_types*d_e];  n_tiles = ceil(n_own/tile_rows), n_chunks = adj_slots/16. */
int tgnn_debug_set_stop_layer(tgnn_handle* h, int32_t layer);
int tgnn_debug_read(tgnn_handle* h, const char* name, float* out, void* stream);
int tgnn_debug_graph(tgnn_handle* h, int32_t* cptr, int32_t* ctype, int32_t* csrc, uint8_t* cdst, float* inv_deg,
                     int32_t* col_ptr, int32_t* col_src, float* type_rows, void* stream);
/* the edge-block format of kernel 3: bptr[ceil(n_own/t_rows)+1] btype[t_blocks] tsrc[t_blocks*128] tdst[t_blocks*128] */
int tgnn_debug_graph_t(tgnn_handle* h, int32_t* bptr, int32_t* btype, int32_t* tsrc, uint16_t* tdst, void* stream);

/* TGNN_ROLE_DBG=1 (environment, read at tgnn_create): per-warp {cycles, wait 0, wait 1, wait 2} of CTA 0 in the last
 * launch of the warp-specialised kernels: out256[0..127] k_conv_t (4 per warp), out256[128..255] k_gin_w. */
int tgnn_debug_role_cycles(tgnn_handle* h, int64_t* out256);

/* Per-kernel-family device time of the last forward (ms), measured with CUDA events on `stream`
 * when enabled.  names: "init","conv","gin","bnfin","combine","final","score","halo". */
int tgnn_set_profiling(tgnn_handle* h, int32_t enabled);
int tgnn_get_profile(tgnn_handle* h, const char* name, float* ms_out, int32_t* launches_out);

const char* tgnn_last_error(tgnn_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* TGNN_H_ */
