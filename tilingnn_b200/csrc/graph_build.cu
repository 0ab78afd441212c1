// Device-side construction of the graph structures the scoring kernels consume.
//
// Input is exactly what TilinGNN.forward receives (graph_networks/networks/TilinGNN.py:51 of the
// reference): int64 COO edge lists in PyG source_to_target order and the fp32 adjacency edge
// features.  Output (struct Graph):
//   * edge-type ids: the reference runs a 3-layer MLP on every edge's feature row
//     (graph_networks/layers/edge_conv.py:17-18); rows that are bitwise equal give equal weights,
//     so edges are labelled with the id of their distinct row (K ids) and the MLP is evaluated K
//     times per layer instead of E_a times.
//   * typed adjacency tiles: destinations are cut into warp tiles of WN rows; the in-edges of a
//     tile are grouped into chunks of CH slots that all share one edge type, and inside every
//     8-slot group all destinations are distinct (so a warp can accumulate messages into shared
//     memory without atomics).
//   * collision CSR by destination with self loops removed (PyG GINConv.remove_self_loops).
// Everything is sorts / scans (CUB) plus small kernels; summation orders derived from it are
// deterministic (stable sorts keyed on content, ties broken by original edge order).
#include <cub/cub.cuh>

#include "tgnn_internal.h"

namespace tgnn {
namespace {

constexpr int TPB = 256;
inline int nblk(int64_t n) { return (int)((n + TPB - 1) / TPB); }

struct MaxOp {
    __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

__global__ void k_validate(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t e,
                           int64_t n_rows, int64_t n_own, int* __restrict__ err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    int64_t s = src[i], d = dst[i];
    if (s < 0 || s >= n_rows || d < 0 || d >= n_own) atomicOr(err, 1);
}

// 64-bit hash of one feature row (-0.0 canonicalised to +0.0 so numerically equal rows match).
__global__ void k_hash_rows(const float* __restrict__ feat, int64_t e, int d_e,
                            unsigned long long* __restrict__ h, int* __restrict__ eid) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    const unsigned* row = reinterpret_cast<const unsigned*>(feat + i * d_e);
    unsigned long long x = 0x9E3779B97F4A7C15ull;
    for (int k = 0; k < d_e; ++k) {
        unsigned w = row[k];
        if (w == 0x80000000u) w = 0u;
        x ^= (unsigned long long)w + 0x9E3779B97F4A7C15ull + (x << 6) + (x >> 2);
        x *= 0xFF51AFD7ED558CCDull;
        x ^= x >> 33;
    }
    h[i] = x;
    eid[i] = (int)i;
}

__global__ void k_head_flags_u64(const unsigned long long* __restrict__ k, int64_t e, int* __restrict__ head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    head[i] = (i == 0 || k[i] != k[i - 1]) ? 1 : 0;
}

// after inclusive scan of head flags: type of sorted position i is scan[i]-1
__global__ void k_assign_types(const int* __restrict__ scan, const int* __restrict__ head,
                               const int* __restrict__ eid_sorted, int64_t e,
                               int* __restrict__ type_of_edge, int* __restrict__ rep_edge, int max_types) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    int t = scan[i] - 1;
    type_of_edge[eid_sorted[i]] = t;
    if (head[i] && t < max_types) rep_edge[t] = eid_sorted[i];
}

__global__ void k_verify_types(const float* __restrict__ feat, int64_t e, int d_e,
                               const int* __restrict__ type_of_edge, const int* __restrict__ rep_edge,
                               int* __restrict__ err) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    const unsigned* a = reinterpret_cast<const unsigned*>(feat + i * d_e);
    const unsigned* b = reinterpret_cast<const unsigned*>(feat + (int64_t)rep_edge[type_of_edge[i]] * d_e);
    bool bad = false;
    for (int k = 0; k < d_e; ++k) {
        unsigned x = a[k], y = b[k];
        if (x == 0x80000000u) x = 0u;
        if (y == 0x80000000u) y = 0u;
        bad |= (x != y);
    }
    if (bad) atomicOr(err, 2);
}

__global__ void k_gather_type_rows(const float* __restrict__ feat, int d_e, const int* __restrict__ rep_edge,
                                   int n_types, float* __restrict__ rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_types * d_e) return;
    int t = i / d_e, k = i - t * d_e;
    rows[i] = feat[(int64_t)rep_edge[t] * d_e + k];
}

// key = (warp tile << (db + tb)) | (type << db) | local destination;  tb = bits of the largest type id, db = log2(wn)
__global__ void k_adj_keys(const int64_t* __restrict__ dst, const int* __restrict__ type_of_edge, int64_t e,
                           int64_t n_own, int tb, int db, unsigned long long* __restrict__ key, int* __restrict__ eid, int* __restrict__ deg) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    long long d = dst[i];
    if (d < 0 || d >= n_own) d = 0;                     // flagged by k_validate; keep the access in range
    unsigned long long tile = (unsigned long long)(d >> db);
    unsigned long long dl = (unsigned long long)(d & ((1ll << db) - 1));
    key[i] = (tile << (db + tb)) | ((unsigned long long)type_of_edge[i] << db) | dl;
    eid[i] = (int)i;
    atomicAdd(&deg[d], 1);
}

__global__ void k_adj_flags(const unsigned long long* __restrict__ key, int64_t e, int db,
                            int* __restrict__ run_head, int* __restrict__ run_start_seed,
                            int* __restrict__ grp_start_seed) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    bool rh = (i == 0) || ((key[i] >> db) != (key[i - 1] >> db));
    bool gh = (i == 0) || (key[i] != key[i - 1]);
    run_head[i] = rh ? 1 : 0;
    run_start_seed[i] = rh ? (int)i : 0;
    grp_start_seed[i] = gh ? (int)i : 0;
}

__global__ void k_adj_runs(const unsigned long long* __restrict__ key, int64_t e,
                           const int* __restrict__ run_idx_incl, const int* __restrict__ run_head,
                           const int* __restrict__ grp_start,
                           int* __restrict__ run_pos, int* __restrict__ run_maxmult) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    int r = run_idx_incl[i] - 1;
    if (run_head[i]) run_pos[r] = (int)i;
    int mult = (int)i - grp_start[i] + 1;
    if (mult > 1) atomicMax(&run_maxmult[r], mult);
}

__global__ void k_run_chunks(const int* __restrict__ run_pos, const int* __restrict__ run_maxmult,
                             int n_runs, int e, int* __restrict__ run_chunks) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    int len = ((r + 1 < n_runs) ? run_pos[r + 1] : e) - run_pos[r];
    int g = (len + GRP - 1) / GRP;
    int mm = run_maxmult[r];
    if (mm > g) g = mm;
    g += g & 1;
    run_chunks[r] = g / 2;
}

// per run: chunk types, and the end of the tile's chunk range if this is the tile's last run
__global__ void k_run_fill(const unsigned long long* __restrict__ key, const int* __restrict__ run_pos,
                           const int* __restrict__ run_chunks, const int* __restrict__ chunk_base,
                           int n_runs, int tb, int db, int* __restrict__ ctype, int* __restrict__ tile_end) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    unsigned long long k = key[run_pos[r]];
    int type = (int)((k >> db) & ((1ull << tb) - 1ull));
    long long tile = (long long)(k >> (db + tb));
    int base = chunk_base[r], nc = run_chunks[r];
    for (int c = 0; c < nc; ++c) ctype[base + c] = type;
    bool last = (r + 1 == n_runs) || ((long long)(key[run_pos[r + 1]] >> (db + tb)) != tile);
    if (last) tile_end[tile] = base + nc;
}

__global__ void k_adj_scatter(const unsigned long long* __restrict__ key, const int* __restrict__ eid_sorted,
                              const int64_t* __restrict__ src, int64_t e,
                              const int* __restrict__ run_idx_incl, const int* __restrict__ run_pos,
                              const int* __restrict__ run_chunks, const int* __restrict__ chunk_base, int db,
                              int* __restrict__ csrc, uint8_t* __restrict__ cdst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    int r = run_idx_incl[i] - 1;
    int p = (int)i - run_pos[r];
    int g = 2 * run_chunks[r];
    int64_t slot = (int64_t)chunk_base[r] * CH + (int64_t)(p % g) * GRP + p / g;
    csrc[slot] = (int)src[eid_sorted[i]];
    cdst[slot] = (uint8_t)(key[i] & ((1ull << db) - 1ull));
}

// Inside every 8-slot group, put destinations of opposite parity next to each other (slots 2i, 2i+1) and lone
// leftovers next to an empty slot.  The two rows a quarter-warp of the adjacency kernels accumulates into (lanes
// g = 2q, 2q+1; row stride 36 floats) then fall on disjoint shared-memory bank groups; rows of equal parity are a 2-way
// conflict on every LDS.128 / STS.128 of the accumulate (ncu: 14 of 98 L1 wavefronts per chunk before this pass).
__global__ void k_pair_parity(int* __restrict__ csrc, uint8_t* __restrict__ cdst, int64_t n_groups) {
    const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    int* ps = csrc + gi * GRP;
    uint8_t* pd = cdst + gi * GRP;
    int s[GRP], d[GRP];
#pragma unroll
    for (int i = 0; i < GRP; ++i) { s[i] = ps[i]; d[i] = pd[i]; }
    int os[GRP], od[GRP];
#pragma unroll
    for (int i = 0; i < GRP; ++i) { os[i] = -1; od[i] = 0; }
    // pass 1: mixed (even, odd) pairs
    int p = 0;
    unsigned used = 0;
    for (;;) {
        int e = -1, o = -1;
#pragma unroll
        for (int i = 0; i < GRP; ++i)
            if (s[i] >= 0 && !((used >> i) & 1u)) {
                if ((d[i] & 1) == 0) { if (e < 0) e = i; }
                else if (o < 0) o = i;
            }
        if (e < 0 || o < 0) break;
        used |= (1u << e) | (1u << o);
#pragma unroll
        for (int i = 0; i < GRP; ++i) {
            if (i == e) { os[p] = s[i]; od[p] = d[i]; }
            if (i == o) { os[p + 1] = s[i]; od[p + 1] = d[i]; }
        }
        p += 2;
    }
    // pass 2: leftovers (all of one parity): first one per remaining pair (its partner stays empty), then the partners
    for (int round = 0; round < 2; ++round)
        for (int q = p + round; q < GRP; q += 2) {
            int pick = -1;
#pragma unroll
            for (int i = 0; i < GRP; ++i)
                if (pick < 0 && s[i] >= 0 && !((used >> i) & 1u)) pick = i;
            if (pick < 0) break;
            used |= 1u << pick;
#pragma unroll
            for (int i = 0; i < GRP; ++i)
                if (i == pick) { os[q] = s[i]; od[q] = d[i]; }
        }
#pragma unroll
    for (int i = 0; i < GRP; ++i) { ps[i] = os[i]; pd[i] = (uint8_t)od[i]; }
}

__global__ void k_fill_int(int* __restrict__ p, int64_t n, int v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_inv_deg(const int* __restrict__ deg, int64_t n, float* __restrict__ inv) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) { int d = deg[i]; inv[i] = 1.0f / (float)(d > 1 ? d : 1); }
}

// cptr[0] = 0, cptr[t+1] = running max of tile_end (empty tiles inherit the previous end)
__global__ void k_cptr_first(int* __restrict__ cptr) { if (threadIdx.x == 0 && blockIdx.x == 0) cptr[0] = 0; }

__global__ void k_col_keys(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t e, int n_own,
                           unsigned* __restrict__ key, int* __restrict__ val, int* __restrict__ cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    long long s = src[i], d = dst[i];
    if (s == d || d < 0 || d >= n_own) { key[i] = (unsigned)n_own; val[i] = -1; }     // self loop: parked behind the last row
    else { key[i] = (unsigned)d; val[i] = (int)s; atomicAdd(&cnt[d], 1); }
}

// ---- S format ------------------------------------------------------------------------------------------
// key = (tile128 << 23) | (type << 7) | local destination row
__global__ void k_s_keys(const int64_t* __restrict__ dst, const int* __restrict__ type_of_edge, int64_t e, int64_t n_own,
                         unsigned long long* __restrict__ key, int* __restrict__ eid) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    long long d = dst[i];
    if (d < 0 || d >= n_own) d = 0;
    key[i] = ((unsigned long long)(d / S_BM) << 23) | ((unsigned long long)type_of_edge[i] << 7) | (unsigned long long)(d % S_BM);
    eid[i] = (int)i;
}
__global__ void k_s_flags(const unsigned long long* __restrict__ key, int64_t e, int* __restrict__ run_head) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    run_head[i] = ((i == 0) || ((key[i] >> 7) != (key[i - 1] >> 7))) ? 1 : 0;
}
__global__ void k_s_heads(const unsigned long long* __restrict__ key, int64_t e, const int* __restrict__ run_idx_incl,
                          const int* __restrict__ run_head, int* __restrict__ pbase, int* __restrict__ ptype,
                          int* __restrict__ tile_end) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    const int pass = run_idx_incl[i] - 1;
    if (run_head[i]) { pbase[pass] = (int)i; ptype[pass] = (int)((key[i] >> 7) & 0xFFFFull); }
    const bool last_of_tile = (i + 1 == e) || ((key[i + 1] >> 23) != (key[i] >> 23));
    if (last_of_tile) tile_end[key[i] >> 23] = pass + 1;
    if (i + 1 == e) pbase[pass + 1] = (int)e;
}
// off[pass][r] = number of edges of the pass whose destination row is < r  (r = 0..128)
__global__ void k_s_offsets(const unsigned long long* __restrict__ key, const int* __restrict__ eid_sorted,
                            const int64_t* __restrict__ src, int64_t e, const int* __restrict__ run_idx_incl,
                            const int* __restrict__ run_head, const int* __restrict__ pbase,
                            unsigned short* __restrict__ off, int* __restrict__ s_src, int* __restrict__ err,
                            int* __restrict__ max_len) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    s_src[i] = (int)src[eid_sorted[i]];
    const int pass = run_idx_incl[i] - 1;
    const int p = (int)i - pbase[pass];
    const int d = (int)(key[i] & 127ull);
    unsigned short* o = off + (size_t)pass * S_OFF_STRIDE;
    const bool grp_head = run_head[i] || (key[i] != key[i - 1]);
    if (grp_head) {
        const int d_prev = run_head[i] ? -1 : (int)(key[i - 1] & 127ull);
        for (int r = d_prev + 1; r <= d; ++r) o[r] = (unsigned short)p;
    }
    const bool last_of_run = (i + 1 == e) || run_head[i + 1];
    if (last_of_run) {
        const int len = p + 1;
        if (len > 65535) atomicOr(err, 4);
        atomicMax(max_len, len);
        for (int r = d + 1; r <= S_BM; ++r) o[r] = (unsigned short)len;
    }
}

// ---- T format (tcgen05 edge-block kernel, conv_t.cu) -----------------------------------------------------------
// Destinations are cut into super-tiles of RT = 2^db rows.  A super-tile's in-edges are grouped into BLOCKS of 128 slots
// that all share one edge type (the A operand of one tcgen05.mma M = 128); quarter q of a block (slots 32q .. 32q+31 =
// TMEM lanes of epilogue warp q) only holds destinations with dst % 4 == q, and inside a quarter all destinations are
// distinct -- so the four epilogue warps accumulate into ONE shared-memory tile without atomics or races.
// key = (tile << (db + tb)) | (type << db) | ((dst & 3) << (db - 2)) | (dst_local >> 2)
__global__ void k_t_keys(const int64_t* __restrict__ dst, const int* __restrict__ type_of_edge, int64_t e, int64_t n_own,
                         int tb, int db, unsigned long long* __restrict__ key, int* __restrict__ eid) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    long long d = dst[i];
    if (d < 0 || d >= n_own) d = 0;
    const unsigned long long tile = (unsigned long long)(d >> db), dl = (unsigned long long)(d & ((1ll << db) - 1));
    key[i] = (tile << (db + tb)) | ((unsigned long long)type_of_edge[i] << db) | ((dl & 3ull) << (db - 2)) | (dl >> 2);
    eid[i] = (int)i;
}
__global__ void k_t_flags(const unsigned long long* __restrict__ key, int64_t e, int db, int* __restrict__ run_head,
                          int* __restrict__ pair_head, int* __restrict__ grp_seed) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    run_head[i] = (i == 0 || (key[i] >> (db - 2)) != (key[i - 1] >> (db - 2))) ? 1 : 0;      // (tile, type, class)
    pair_head[i] = (i == 0 || (key[i] >> db) != (key[i - 1] >> db)) ? 1 : 0;                  // (tile, type)
    grp_seed[i] = (i == 0 || key[i] != key[i - 1]) ? (int)i : 0;                              // same destination
}
__global__ void k_t_runs(int64_t e, const int* __restrict__ run_idx, const int* __restrict__ run_head,
                         const int* __restrict__ pair_idx, const int* __restrict__ pair_head, const int* __restrict__ grp_start,
                         int* __restrict__ run_pos, int* __restrict__ run_pair, int* __restrict__ run_maxmult, int* __restrict__ pair_pos) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    const int r = run_idx[i] - 1;
    if (run_head[i]) { run_pos[r] = (int)i; run_pair[r] = pair_idx[i] - 1; }
    if (pair_head[i]) pair_pos[pair_idx[i] - 1] = (int)i;
    const int mult = (int)i - grp_start[i] + 1;
    if (mult > 1) atomicMax(&run_maxmult[r], mult);
}
// groups (quarter-blocks of 32 slots) per run; blocks per (tile, type) pair = the most groups any of its classes needs
__global__ void k_t_groups(const int* __restrict__ run_pos, const int* __restrict__ run_pair, const int* __restrict__ run_maxmult,
                           int n_runs, int e, int* __restrict__ run_g, int* __restrict__ pair_nb) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const int len = ((r + 1 < n_runs) ? run_pos[r + 1] : e) - run_pos[r];
    int g = (len + 31) / 32;
    if (run_maxmult[r] > g) g = run_maxmult[r];
    run_g[r] = g;
    atomicMax(&pair_nb[run_pair[r]], g);
}
__global__ void k_t_pairs(const unsigned long long* __restrict__ key, const int* __restrict__ pair_pos, const int* __restrict__ pair_nb,
                          const int* __restrict__ pair_base, int n_pairs, int tb, int db, int* __restrict__ tile_end) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const long long tile = (long long)(key[pair_pos[p]] >> (db + tb));
    const bool last = (p + 1 == n_pairs) || ((long long)(key[pair_pos[p + 1]] >> (db + tb)) != tile);
    if (last) tile_end[tile] = pair_base[p] + pair_nb[p];             // edge blocks of all tiles up to and including this one
}
__global__ void k_t_bptr(const int* __restrict__ cum, int n_tiles, int rb, int* __restrict__ bptr) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) bptr[0] = 0;
    if (t < n_tiles) bptr[t + 1] = cum[t] + rb * (t + 1);             // + the root blocks of every tile so far
}
__global__ void k_t_btype(const unsigned long long* __restrict__ key, const int* __restrict__ pair_pos, const int* __restrict__ pair_nb,
                          const int* __restrict__ pair_base, int n_pairs, int tb, int db, int rb, int* __restrict__ btype) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const unsigned long long k = key[pair_pos[p]];
    const int type = (int)((k >> db) & ((1ull << tb) - 1ull));
    const int tile = (int)(k >> (db + tb));
    for (int b = 0; b < pair_nb[p]; ++b) btype[pair_base[p] + rb * tile + b] = type;
}
__global__ void k_t_scatter(const unsigned long long* __restrict__ key, const int* __restrict__ eid_sorted, const int64_t* __restrict__ src,
                            int64_t e, const int* __restrict__ run_idx, const int* __restrict__ run_pos, const int* __restrict__ run_g,
                            const int* __restrict__ run_pair, const int* __restrict__ pair_base, int tb, int db, int rb,
                            int* __restrict__ t_src, unsigned short* __restrict__ t_dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= e) return;
    const int r = run_idx[i] - 1, p = (int)i - run_pos[r], g = run_g[r];
    const unsigned long long k = key[i];
    const int tile = (int)(k >> (db + tb)), q = (int)((k >> (db - 2)) & 3ull);
    const int dl = (int)((k & ((1ull << (db - 2)) - 1ull)) << 2) | q;
    const int64_t slot = ((int64_t)pair_base[run_pair[r]] + (int64_t)rb * tile + (p % g)) * 128 + 32 * q + p / g;
    t_src[slot] = (int)src[eid_sorted[i]];
    t_dst[slot] = (unsigned short)dl;
}
// root blocks (the x_i . root term as rb more blocks of edge type n_types, src = dst = the tile's own rows)
__global__ void k_t_roots(const int* __restrict__ bptr, int n_tiles, int rt, int rb, int n_own, int n_types,
                          int* __restrict__ btype, int* __restrict__ t_src, unsigned short* __restrict__ t_dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_tiles * rt) return;
    const int tile = (int)(i / rt), r = (int)(i % rt);
    const int block = bptr[tile + 1] - rb + (r >> 7), rr = r & 127;
    const int64_t slot = (int64_t)block * 128 + 32 * (rr & 3) + (rr >> 2);
    const int node = tile * rt + r;
    if (node < n_own) { t_src[slot] = node; t_dst[slot] = (unsigned short)r; }
    if (rr == 0) btype[block] = n_types;
}
__global__ void k_fill_u16(unsigned short* __restrict__ p, int64_t n, unsigned short v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int bits_for(unsigned long long v) { int b = 1; while ((v >> b) != 0 && b < 64) ++b; return b; }

template <class K, class V>
void sort_pairs(Scratch& sc, const K* kin, K* kout, const V* vin, V* vout, int64_t n, int end_bit, cudaStream_t st) {
    size_t tb = 0;
    TGNN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, (int)n, 0, end_bit, st));
    void* tmp = sc.get<char>(tb);
    TGNN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, kin, kout, vin, vout, (int)n, 0, end_bit, st));
}

void incl_sum(Scratch& sc, const int* in, int* out, int64_t n, cudaStream_t st) {
    size_t tb = 0;
    TGNN_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, in, out, (int)n, st));
    void* tmp = sc.get<char>(tb);
    TGNN_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, in, out, (int)n, st));
}
void excl_sum(Scratch& sc, const int* in, int* out, int64_t n, cudaStream_t st) {
    size_t tb = 0;
    TGNN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int)n, st));
    void* tmp = sc.get<char>(tb);
    TGNN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, (int)n, st));
}
void incl_max(Scratch& sc, const int* in, int* out, int64_t n, cudaStream_t st) {
    size_t tb = 0;
    TGNN_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, in, out, MaxOp(), (int)n, st));
    void* tmp = sc.get<char>(tb);
    TGNN_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb, in, out, MaxOp(), (int)n, st));
}

int read_int(const int* dptr, cudaStream_t st) {
    int v = 0;
    TGNN_CUDA(cudaMemcpyAsync(&v, dptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    TGNN_CUDA(cudaStreamSynchronize(st));
    return v;
}

}  // namespace

void build_graph(Graph& g, Scratch& sc, int d_e, int64_t n_own, int64_t n_rows,
                 int64_t e_adj, const int64_t* adj_src, const int64_t* adj_dst, const float* adj_feat,
                 int64_t e_col, const int64_t* col_src, const int64_t* col_dst, int want_s, int wn, int want_t, cudaStream_t st) {
    TGNN_CHECK(wn == WN_SMALL || wn == WN_BIG, "internal: bad warp-tile height");
    const int db = wn == WN_BIG ? 7 : 6;
    g.wn = wn;
    TGNN_CHECK(n_own > 0 && n_rows >= n_own, "tgnn_set_graph: need n_nodes > 0");
    TGNN_CHECK(n_rows < (1ll << 31) - 64, "tgnn_set_graph: more than 2^31 rows per GPU is not supported");
    TGNN_CHECK(e_adj >= 0 && e_adj < (1ll << 31) - 64 && e_col >= 0 && e_col < (1ll << 31) - 64,
               "tgnn_set_graph: edge count per GPU must be below 2^31");
    sc.reset();
    g.n_own = n_own; g.n_rows = n_rows;
    g.n_tiles = (int)((n_own + wn - 1) / wn);
    int* err = sc.get<int>(1);
    TGNN_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));

    // ---------------- adjacency: edge types -----------------------------------------------------
    g.cptr.reserve((size_t)(g.n_tiles + 1) * sizeof(int));
    g.inv_deg.reserve((size_t)n_own * sizeof(float));
    int* deg = sc.get<int>(n_own);
    TGNN_CUDA(cudaMemsetAsync(deg, 0, (size_t)n_own * sizeof(int), st));
    g.n_types = 0; g.n_chunks = 0; g.e_adj = e_adj;
    if (e_adj > 0) {
        k_validate<<<nblk(e_adj), TPB, 0, st>>>(adj_src, adj_dst, e_adj, n_rows, n_own, err);
        unsigned long long* h0 = sc.get<unsigned long long>(e_adj);
        unsigned long long* h1 = sc.get<unsigned long long>(e_adj);
        int* id0 = sc.get<int>(e_adj);
        int* id1 = sc.get<int>(e_adj);
        k_hash_rows<<<nblk(e_adj), TPB, 0, st>>>(adj_feat, e_adj, d_e, h0, id0);
        sort_pairs(sc, h0, h1, id0, id1, e_adj, 64, st);
        int* head = sc.get<int>(e_adj);
        int* scan = sc.get<int>(e_adj);
        k_head_flags_u64<<<nblk(e_adj), TPB, 0, st>>>(h1, e_adj, head);
        incl_sum(sc, head, scan, e_adj, st);
        int n_types = read_int(scan + (e_adj - 1), st);
        TGNN_CHECK(read_int(err, st) == 0, "tgnn_set_graph: adjacency edge index out of range");
        TGNN_CHECK(n_types <= MAX_TYPES,
                   "tgnn_set_graph: more than 4194304 distinct adjacency edge-feature rows (one 12 KB weight table each "
                   "per layer would not fit the GPU)");
        g.n_types = n_types;
        int* type_of_edge = sc.get<int>(e_adj);
        int* rep_edge = sc.get<int>(n_types);
        k_assign_types<<<nblk(e_adj), TPB, 0, st>>>(scan, head, id1, e_adj, type_of_edge, rep_edge, n_types);
        k_verify_types<<<nblk(e_adj), TPB, 0, st>>>(adj_feat, e_adj, d_e, type_of_edge, rep_edge, err);
        g.type_rows.reserve((size_t)n_types * d_e * sizeof(float));
        k_gather_type_rows<<<nblk((int64_t)n_types * d_e), TPB, 0, st>>>(adj_feat, d_e, rep_edge, n_types,
                                                                         g.type_rows.as<float>());

        // ---------------- adjacency: typed tiles ---------------------------------------------------
        unsigned long long* k0 = h0;      // reuse
        unsigned long long* k1 = h1;
        const int tb = bits_for((unsigned long long)(n_types > 0 ? n_types - 1 : 0));
        k_adj_keys<<<nblk(e_adj), TPB, 0, st>>>(adj_dst, type_of_edge, e_adj, n_own, tb, db, k0, id0, deg);
        int end_bit = db + tb + bits_for((unsigned long long)g.n_tiles);
        TGNN_CHECK(end_bit <= 64, "tgnn_set_graph: sort key overflow (tiles x edge types)");
        sort_pairs(sc, k0, k1, id0, id1, e_adj, end_bit, st);
        int* run_head = head;
        int* run_seed = sc.get<int>(e_adj);
        int* grp_seed = sc.get<int>(e_adj);
        k_adj_flags<<<nblk(e_adj), TPB, 0, st>>>(k1, e_adj, db, run_head, run_seed, grp_seed);
        int* run_idx = scan;
        incl_sum(sc, run_head, run_idx, e_adj, st);
        int* grp_start = sc.get<int>(e_adj);
        incl_max(sc, grp_seed, grp_start, e_adj, st);
        int n_runs = read_int(run_idx + (e_adj - 1), st);
        int* run_pos = sc.get<int>(n_runs + 1);
        int* run_maxmult = sc.get<int>(n_runs);
        int* run_chunks = sc.get<int>(n_runs);
        int* chunk_base = sc.get<int>(n_runs + 1);
        TGNN_CUDA(cudaMemsetAsync(run_maxmult, 0, (size_t)n_runs * sizeof(int), st));
        k_adj_runs<<<nblk(e_adj), TPB, 0, st>>>(k1, e_adj, run_idx, run_head, grp_start, run_pos, run_maxmult);
        k_run_chunks<<<nblk(n_runs), TPB, 0, st>>>(run_pos, run_maxmult, n_runs, (int)e_adj, run_chunks);
        excl_sum(sc, run_chunks, chunk_base, n_runs, st);
        int last_base = read_int(chunk_base + (n_runs - 1), st);
        int last_cnt = read_int(run_chunks + (n_runs - 1), st);
        int64_t n_chunks = (int64_t)last_base + last_cnt;
        TGNN_CHECK(n_chunks * CH < (1ll << 31) - 64, "tgnn_set_graph: typed adjacency tiles exceed 2^31 slots");
        g.n_chunks = (int)n_chunks;
        g.ctype.reserve((size_t)n_chunks * sizeof(int));
        g.csrc.reserve((size_t)n_chunks * CH * sizeof(int));
        g.cdst.reserve((size_t)n_chunks * CH);
        k_fill_int<<<nblk(n_chunks * CH), TPB, 0, st>>>(g.csrc.as<int>(), n_chunks * CH, -1);
        TGNN_CUDA(cudaMemsetAsync(g.cdst.p, 0, (size_t)n_chunks * CH, st));
        int* tile_end = sc.get<int>(g.n_tiles);
        TGNN_CUDA(cudaMemsetAsync(tile_end, 0, (size_t)g.n_tiles * sizeof(int), st));
        k_run_fill<<<nblk(n_runs), TPB, 0, st>>>(k1, run_pos, run_chunks, chunk_base, n_runs, tb, db,
                                                  g.ctype.as<int>(), tile_end);
        k_cptr_first<<<1, 32, 0, st>>>(g.cptr.as<int>());
        incl_max(sc, tile_end, g.cptr.as<int>() + 1, g.n_tiles, st);
        k_adj_scatter<<<nblk(e_adj), TPB, 0, st>>>(k1, id1, adj_src, e_adj, run_idx, run_pos, run_chunks,
                                                    chunk_base, db, g.csrc.as<int>(), g.cdst.as<uint8_t>());
        k_pair_parity<<<nblk(n_chunks * (CH / GRP)), TPB, 0, st>>>(g.csrc.as<int>(), g.cdst.as<uint8_t>(), n_chunks * (CH / GRP));
        // ---------------- adjacency: S format (tcgen05 kernel) -----------------------------------------
        g.has_s = false; g.s_built = false; g.has_z = false;
        // the S format costs a 64-bit sort; in "auto" mode (want_s == 1) it is only built when the tcgen05 kernel could be
        // chosen at all: it needs ~164 edges per (128-row tile, type) pass to beat the fp16 edge-chunk kernel
        const double passes_est = 0.8 * (double)((n_own + S_BM - 1) / S_BM) *
                                  std::min<double>((double)n_types, (double)e_adj * S_BM / (double)n_own);
        const bool s_may_win = want_s >= 2 || passes_est * S_EDGES_PER_PASS_BREAK_EVEN < (double)e_adj;
        if (want_s && s_may_win && n_types <= S_MAX_TYPES) {
            g.s_tiles = (int)((n_own + S_BM - 1) / S_BM);
            k_s_keys<<<nblk(e_adj), TPB, 0, st>>>(adj_dst, type_of_edge, e_adj, n_own, k0, id0);
            sort_pairs(sc, k0, k1, id0, id1, e_adj, 23 + bits_for((unsigned long long)g.s_tiles), st);
            k_s_flags<<<nblk(e_adj), TPB, 0, st>>>(k1, e_adj, run_head);
            incl_sum(sc, run_head, run_idx, e_adj, st);
            const int n_pass = read_int(run_idx + (e_adj - 1), st);
            g.s_passes = n_pass;
            g.s_pptr.reserve((size_t)(g.s_tiles + 1) * sizeof(int));
            g.s_ptype.reserve((size_t)n_pass * sizeof(int));
            g.s_pbase.reserve((size_t)(n_pass + 1) * sizeof(int));
            g.s_off.reserve((size_t)n_pass * S_OFF_STRIDE * sizeof(unsigned short));
            g.s_src.reserve((size_t)e_adj * sizeof(int));
            int* s_tile_end = sc.get<int>(g.s_tiles);
            TGNN_CUDA(cudaMemsetAsync(s_tile_end, 0, (size_t)g.s_tiles * sizeof(int), st));
            k_s_heads<<<nblk(e_adj), TPB, 0, st>>>(k1, e_adj, run_idx, run_head, g.s_pbase.as<int>(), g.s_ptype.as<int>(), s_tile_end);
            int* s_maxlen = sc.get<int>(1);
            TGNN_CUDA(cudaMemsetAsync(s_maxlen, 0, sizeof(int), st));
            k_s_offsets<<<nblk(e_adj), TPB, 0, st>>>(k1, id1, adj_src, e_adj, run_idx, run_head, g.s_pbase.as<int>(),
                                                      g.s_off.as<unsigned short>(), g.s_src.as<int>(), err, s_maxlen);
            k_cptr_first<<<1, 32, 0, st>>>(g.s_pptr.as<int>());
            incl_max(sc, s_tile_end, g.s_pptr.as<int>() + 1, g.s_tiles, st);
            g.s_max_pass = read_int(s_maxlen, st);
            g.s_built = true;
            g.has_s = g.s_max_pass <= 160;          // four passes in flight must fit the kernel's 640-row ring
        }
        // ---------------- adjacency: T format (tcgen05 edge-block kernel) --------------------------------
        g.has_t = false;
        if (want_t > 0) {
            const int rt = want_t, tdb = rt == 512 ? 9 : 8, rb = rt / 128;
            TGNN_CHECK(rt == 256 || rt == 512, "internal: bad super-tile height");
            g.t_rows = rt; g.t_tiles = (int)((n_own + rt - 1) / rt);
            k_t_keys<<<nblk(e_adj), TPB, 0, st>>>(adj_dst, type_of_edge, e_adj, n_own, tb, tdb, k0, id0);
            const int t_end_bit = tdb + tb + bits_for((unsigned long long)g.t_tiles);
            TGNN_CHECK(t_end_bit <= 64, "tgnn_set_graph: sort key overflow (super-tiles x edge types)");
            sort_pairs(sc, k0, k1, id0, id1, e_adj, t_end_bit, st);
            int* pair_head = sc.get<int>(e_adj);
            int* pair_idx = sc.get<int>(e_adj);
            k_t_flags<<<nblk(e_adj), TPB, 0, st>>>(k1, e_adj, tdb, run_head, pair_head, grp_seed);
            incl_sum(sc, run_head, run_idx, e_adj, st);
            incl_sum(sc, pair_head, pair_idx, e_adj, st);
            incl_max(sc, grp_seed, grp_start, e_adj, st);
            const int t_runs = read_int(run_idx + (e_adj - 1), st), t_pairs = read_int(pair_idx + (e_adj - 1), st);
            int* t_run_pos = sc.get<int>(t_runs + 1);
            int* t_run_pair = sc.get<int>(t_runs);
            int* t_run_mm = sc.get<int>(t_runs);
            int* t_run_g = sc.get<int>(t_runs);
            int* pair_pos = sc.get<int>(t_pairs + 1);
            int* pair_nb = sc.get<int>(t_pairs);
            int* pair_base = sc.get<int>(t_pairs + 1);
            TGNN_CUDA(cudaMemsetAsync(t_run_mm, 0, (size_t)t_runs * sizeof(int), st));
            TGNN_CUDA(cudaMemsetAsync(pair_nb, 0, (size_t)t_pairs * sizeof(int), st));
            k_t_runs<<<nblk(e_adj), TPB, 0, st>>>(e_adj, run_idx, run_head, pair_idx, pair_head, grp_start, t_run_pos, t_run_pair, t_run_mm, pair_pos);
            k_t_groups<<<nblk(t_runs), TPB, 0, st>>>(t_run_pos, t_run_pair, t_run_mm, t_runs, (int)e_adj, t_run_g, pair_nb);
            excl_sum(sc, pair_nb, pair_base, t_pairs, st);
            const int64_t edge_blocks = (int64_t)read_int(pair_base + (t_pairs - 1), st) + read_int(pair_nb + (t_pairs - 1), st);
            const int64_t n_blocks = edge_blocks + (int64_t)rb * g.t_tiles;
            TGNN_CHECK(n_blocks * 128 < (1ll << 31) - 64, "tgnn_set_graph: edge blocks exceed 2^31 slots");
            g.t_blocks = (int)n_blocks;
            g.t_bptr.reserve((size_t)(g.t_tiles + 1) * sizeof(int));
            g.t_btype.reserve((size_t)n_blocks * sizeof(int));
            g.t_src.reserve((size_t)n_blocks * 128 * sizeof(int));
            g.t_dst.reserve((size_t)n_blocks * 128 * sizeof(unsigned short));
            k_fill_int<<<nblk(n_blocks * 128), TPB, 0, st>>>(g.t_src.as<int>(), n_blocks * 128, -1);
            k_fill_u16<<<nblk(n_blocks * 128), TPB, 0, st>>>(g.t_dst.as<unsigned short>(), n_blocks * 128, (unsigned short)0xFFFF);
            int* t_tile_end = sc.get<int>(g.t_tiles);
            int* t_cum = sc.get<int>(g.t_tiles);
            TGNN_CUDA(cudaMemsetAsync(t_tile_end, 0, (size_t)g.t_tiles * sizeof(int), st));
            k_t_pairs<<<nblk(t_pairs), TPB, 0, st>>>(k1, pair_pos, pair_nb, pair_base, t_pairs, tb, tdb, t_tile_end);
            incl_max(sc, t_tile_end, t_cum, g.t_tiles, st);
            k_t_bptr<<<nblk(g.t_tiles), TPB, 0, st>>>(t_cum, g.t_tiles, rb, g.t_bptr.as<int>());
            k_t_btype<<<nblk(t_pairs), TPB, 0, st>>>(k1, pair_pos, pair_nb, pair_base, t_pairs, tb, tdb, rb, g.t_btype.as<int>());
            k_t_scatter<<<nblk(e_adj), TPB, 0, st>>>(k1, id1, adj_src, e_adj, run_idx, t_run_pos, t_run_g, t_run_pair, pair_base, tb, tdb, rb,
                                                      g.t_src.as<int>(), g.t_dst.as<unsigned short>());
            k_t_roots<<<nblk((int64_t)g.t_tiles * rt), TPB, 0, st>>>(g.t_bptr.as<int>(), g.t_tiles, rt, rb, (int)n_own, n_types,
                                                                    g.t_btype.as<int>(), g.t_src.as<int>(), g.t_dst.as<unsigned short>());
            g.has_t = true;
        }
    } else {
        g.has_s = false; g.has_t = false; g.s_built = false; g.has_z = false;
        TGNN_CUDA(cudaMemsetAsync(g.cptr.p, 0, (size_t)(g.n_tiles + 1) * sizeof(int), st));
    }
    k_inv_deg<<<nblk(n_own), TPB, 0, st>>>(deg, n_own, g.inv_deg.as<float>());

    // ---------------- collision CSR -----------------------------------------------------------------
    g.col_ptr.reserve((size_t)(n_own + 1 + 72) * sizeof(int));     // + padding: k_gin_w copies 68 pointers per 64-row tile
    g.e_col = 0;
    if (e_col > 0) {
        k_validate<<<nblk(e_col), TPB, 0, st>>>(col_src, col_dst, e_col, n_rows, n_own, err);
        unsigned* ck0 = sc.get<unsigned>(e_col);
        unsigned* ck1 = sc.get<unsigned>(e_col);
        int* cv0 = sc.get<int>(e_col);
        g.col_src.reserve((size_t)e_col * sizeof(int));
        int* cnt = sc.get<int>(n_own + 1);
        TGNN_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(n_own + 1) * sizeof(int), st));
        k_col_keys<<<nblk(e_col), TPB, 0, st>>>(col_src, col_dst, e_col, (int)n_own, ck0, cv0, cnt);
        sort_pairs(sc, ck0, ck1, cv0, g.col_src.as<int>(), e_col, bits_for((unsigned long long)n_own), st);
        excl_sum(sc, cnt, g.col_ptr.as<int>(), n_own + 1, st);
        g.e_col = read_int(g.col_ptr.as<int>() + n_own, st);
    } else {
        TGNN_CUDA(cudaMemsetAsync(g.col_ptr.p, 0, (size_t)(n_own + 1) * sizeof(int), st));
    }
    int e = read_int(err, st);
    TGNN_CHECK((e & 1) == 0, "tgnn_set_graph: edge index out of range");
    TGNN_CHECK((e & 2) == 0, "tgnn_set_graph: 64-bit hash collision between distinct edge-feature rows");
    TGNN_CHECK((e & 4) == 0, "tgnn_set_graph: more than 65535 same-type in-edges in one 128-row tile");
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
