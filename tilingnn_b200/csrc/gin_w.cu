// Collision branch with STAGED NEIGHBOUR WINDOWS ("W" kernel): GINConv sum + node MLP + LeakyReLU + BatchNorm partial
// sums (graph_networks/layers/coll_conv.py:24-27 of the reference; PyG GINConv), same arithmetic as k_gin (kernels.cu).
//
// What changes is where the neighbour rows come from.  k_gin gathers every edge's 128-byte row with a per-lane global
// load: 4 GB of L2 -> SM traffic per launch at 1M nodes x deg 32 and a kernel that waits on L2 latency (ncu: 3.3 of
// 5.7 stall cycles per issue are long-scoreboard, 24 % occupancy).  On tile graphs the in-edges of 64 consecutive
// destinations touch a few hundred DISTINCT source rows in a handful of contiguous runs (lattice, deg 32: 2112 edges,
// 648 rows, 9 runs).  graph_build.cu therefore emits, per 64-row tile, the run list ("window") and rewrites the CSR
// sources as uint16 window-local offsets; here a producer warp brings the window into shared memory with TMA bulk
// copies (cp.async.bulk -> UBLKCP, completion on an mbarrier, double buffered against the previous tile's compute)
// and the gather warps read neighbour rows with LDS.128 at shared-memory latency:
//   L2 -> SM traffic / 3.3, no dependent global load in the gather loop, indices 2 B instead of 4 B per edge.
// Warp roles (16 warps = 512 threads x 128 registers, one CTA per SM, persistent over tiles):
//   warps 0-6   node MLP on tensor cores (mma.sync): the four 16-node chunks of a tile go round-robin over the 7 warps
//               (each has its own chunk buffer and barrier pair), so ~two tiles' MLPs are in flight (one chunk is ~6000 cycles
//               of dependent MMAs and sigmoids)
//   warps 7-14  gather-sum from the window (8 lanes per 128-byte row, 16 rows in flight per lane; quads of 4 destination rows
//               round-robin over the eight warps).  This is the role that bounds the kernel.  The split is measured
//               (TGNN_ROLE_DBG, profiles/r2/gin_w_role_cycles.txt): with 4 gather / 11 MLP warps the gather warps were busy 90 %
//               of the time and the MLP warps waited 76 % of it (the MLP had meanwhile lost a third of its instructions: SFU
//               sigmoid, packed fp32 adds); 6 / 9: 0.377, 8 / 7: 0.334, 9 / 6: 0.340, 10 / 5: 0.352 ms per launch (4 / 11: 0.393)
//   warp  15    producer: per tile <= 32 bulk copies (window runs), 1 for the uint16 indices, 1 for the row pointers
// Tiles whose sources are not local (window > 656 rows / > 32 runs / > 4032 edges) are marked "direct" by the
// builder and gathered from global memory by the same warps; graphs that are mostly direct keep k_gin.
#include <cub/cub.cuh>

#include "gin_mlp.cuh"
#include "tc_common.cuh"
#include "tgnn_internal.h"

namespace tgnn {
namespace {

using namespace ginx;
using namespace tc;

constexpr int GATHER_WARPS = 15 - GW_MLP_WARPS, MLP_WARPS = GW_MLP_WARPS, CHUNKS_PER_TILE = GW_T / CH, QUADS_PER_TILE = GW_T / 4;
constexpr int W_GATHER0 = MLP_WARPS, W_PROD = MLP_WARPS + GATHER_WARPS;
constexpr int GW_THREADS = (W_PROD + 1) * 32;
constexpr int PTR_INTS = 68;                                    // 65 row pointers, padded to a multiple of 16 bytes
constexpr int OFF_WIN = 0;                                      // [2][GW_WMAX][128 B]
constexpr int OFF_LOC = OFF_WIN + 2 * GW_WMAX * 128;            // [2][GW_CAP] uint16
constexpr int OFF_PTR = OFF_LOC + 2 * GW_CAP * 2;               // [2][PTR_INTS] int
constexpr int OFF_S = OFF_PTR + 2 * PTR_INTS * 4;               // [MLP_WARPS][CH][XS] float: one 16-node chunk buffer per MLP warp
constexpr int OFF_W = OFF_S + MLP_WARPS * CH * XS * 4;          // MLP weights
template <bool HMLP> constexpr int gw_smem_bytes() { return OFF_W + (HMLP ? GIN_WFLOATS_H : GIN_WFLOATS) * 4 + 128; }

__device__ __forceinline__ float4 lds_row(uint32_t addr) { return lds128f(addr); }

template <bool HMLP>
__global__ void __launch_bounds__(GW_THREADS, 1)
k_gin_w(GinArgs A) {
    // (no manual alignment of the dynamic shared memory: a pointer that went through an integer cast loses its address
    // space and every access through it becomes a generic 64-bit LD/ST -- measured: the gather warps were bound by exactly
    // that address arithmetic.  16-byte alignment is all the bulk copies and the 128-bit accesses need.)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4 + 2 * MLP_WARPS];   // win_full[2], win_empty[2], chunk_full[MLP_WARPS], chunk_empty[MLP_WARPS]
    __shared__ int meta_s[2][4];                                // per window buffer: {nseg (0 = direct), self_loc, -, -}
    __shared__ int timeout_flag;
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_wf = smem_u32(&bars[0]), bar_we = smem_u32(&bars[2]), bar_cf = smem_u32(&bars[4]), bar_ce = smem_u32(&bars[4 + MLP_WARPS]);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar_wf + 8 * i, 1); mbar_init(bar_we + 8 * i, GATHER_WARPS); }
        // chunk ci of the CTA (ci = 4 * tile iteration + chunk in tile) belongs to MLP warp ci % MLP_WARPS and is that warp's
        // (ci / MLP_WARPS)-th chunk: every party of a chunk barrier sees every one of its phases (a parity wait must not skip one)
        for (int i = 0; i < MLP_WARPS; ++i) { mbar_init(bar_cf + 8 * i, CH / 4); mbar_init(bar_ce + 8 * i, 1); }       // one arrival per quad of the chunk
        timeout_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float* wsm = reinterpret_cast<float*>(smem + OFF_W);
    gin_load_weights<HMLP>(wsm, A.wfrag, tid, GW_THREADS);
    __syncthreads();
    const int n_tiles = A.gw_tiles;
    long long w0 = 0, w1 = 0, w2 = 0;
    const long long t_start = clock64();

    if (warp == W_PROD) {
        // ===================== producer: window runs, indices and row pointers by TMA bulk copy =====================
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            if (!TGNN_TIMED(w0, mbar_wait_relaxed(bar_we + 8 * b, (uint32_t)(((it >> 1) & 1) ^ 1)))) { timeout_flag = 1; break; }
            const int4 m0 = __ldg(reinterpret_cast<const int4*>(A.gw_meta) + 2 * tile);       // {nseg, rows, loc offset, self_loc}
            const int ne = __ldg(A.gw_meta + 8 * tile + 4);
            const int nseg = m0.x, rows = m0.y;
            const int ne_pad = (ne + 7) & ~7;
            const uint32_t bar = bar_wf + 8 * b;
            if (lane == 0) {
                meta_s[b][0] = nseg; meta_s[b][1] = m0.w;
                const uint32_t bytes = (uint32_t)PTR_INTS * 4u + (nseg > 0 ? (uint32_t)rows * 128u + (uint32_t)ne_pad * 2u : 0u);
                mbar_arrive_expect_tx(bar, bytes);               // (release: the meta_s stores above are visible to the waiters)
            }
            __syncwarp();
            if (lane < nseg) {
                const int2 sg = __ldg(reinterpret_cast<const int2*>(A.gw_seg) + (size_t)tile * GW_MAXSEG + lane);
                const int next = lane + 1 < nseg ? __ldg(A.gw_seg + ((size_t)tile * GW_MAXSEG + lane + 1) * 2 + 1) : rows;
                bulk_g2s(sbase + OFF_WIN + (uint32_t)b * (GW_WMAX * 128) + (uint32_t)sg.y * 128u,
                         A.xin + (size_t)sg.x * F, (uint32_t)(next - sg.y) * 128u, bar);
            }
            if (lane == 0) {
                if (nseg > 0 && ne_pad > 0)
                    bulk_g2s(sbase + OFF_LOC + (uint32_t)b * (GW_CAP * 2), A.gw_loc + m0.z, (uint32_t)ne_pad * 2u, bar);
                bulk_g2s(sbase + OFF_PTR + (uint32_t)b * (PTR_INTS * 4), A.col_ptr + (size_t)tile * GW_T, (uint32_t)PTR_INTS * 4u, bar);
            }
        }
    } else if (warp >= W_GATHER0) {
        // ===================== gather-sum: 8 lanes per row, each warp two quads of 4 destination rows =====================
        const int gw = warp - W_GATHER0;
        const int a = lane >> 3, q = lane & 7;
        const float self_w = 1.0f + A.eps;
        int rk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) rk[k] = (k + 2 * a) & 7;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            if (!TGNN_TIMED(w0, mbar_wait(bar_wf + 8 * b, ph))) { timeout_flag = 1; break; }
            const int nseg = meta_s[b][0], self_loc = meta_s[b][1];
            const int* ptr = reinterpret_cast<const int*>(smem + OFF_PTR + b * (PTR_INTS * 4));
            const uint32_t loc_s = sbase + OFF_LOC + (uint32_t)b * (GW_CAP * 2);
            const uint32_t win = sbase + OFF_WIN + (uint32_t)b * (GW_WMAX * 128) + (uint32_t)q * 16u;
            const int node0 = tile * GW_T, e_base = ptr[0];
            bool ok = true;
#pragma unroll 1
            for (int qd = 0; qd < QUADS_PER_TILE; ++qd) {
                // quads (4 destination rows, one per 8-lane group) go round-robin over the gather warps ACROSS tiles; a quad's
                // sums go into the chunk buffer of the MLP warp that owns its chunk
                if ((it * QUADS_PER_TILE + qd) % GATHER_WARPS != gw) continue;
                const int c = qd >> 2, ci = it * CHUNKS_PER_TILE + c, mw = ci % MLP_WARPS, use = ci / MLP_WARPS;
                float* S = reinterpret_cast<float*>(smem + OFF_S) + mw * (CH * XS);
                const int r = 4 * qd + a, node = node0 + r;
                const bool live = node < A.n_own;
                const int e0 = live ? ptr[r] : e_base, n_mine = live ? ptr[r + 1] - e0 : 0;
                int n_max = max(n_mine, __shfl_xor_sync(0xffffffffu, n_mine, 8));       // warp-uniform trip count
                n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, 16));
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                if (nseg > 0) {
                    if (live) {
                        const float4 c = lds_row(win + (uint32_t)(self_loc + r) * 128u);
                        sum.x = self_w * c.x; sum.y = self_w * c.y; sum.z = self_w * c.z; sum.w = self_w * c.w;
                    }
                    // 16 rows in flight per lane (one gather warp per scheduler: the loads of a batch must cover the
                    // LDS -> address -> LDS.128 -> FADD chain of the previous one); same summation order as k_gin: batches of
                    // 8, each walked in the lane group's rotated order (k + 2a) -- with equal degrees the four groups' index
                    // lists are a fixed distance apart and the same position would fall on one bank.  Positions past the
                    // row's last neighbour read a stale index (still inside the staging buffer) that is never used.
                    const uint32_t lp = loc_s + 2u * (uint32_t)(e0 - e_base);
                    for (int o = 0; o < n_max; o += 16) {
                        uint32_t ad[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) ad[k] = win + lds_u16(lp + 2u * (uint32_t)(o + (k & 8) + rk[k & 7])) * 128u;
                        float4 v[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (o + (k & 8) + rk[k & 7] < n_mine) v[k] = lds_row(ad[k]);
                        }
#pragma unroll
                        for (int k = 0; k < 16; ++k) f4add(sum, v[k]);          // two FADD2 per row
                    }
                } else {
                    // "direct" tile (sources not local enough for a window): rows and indices from global memory
                    if (live) {
                        const float4 c = ld_row4(A.xin, node, q);
                        sum.x = self_w * c.x; sum.y = self_w * c.y; sum.z = self_w * c.z; sum.w = self_w * c.w;
                    }
                    const int last = max(n_mine - 1, 0);
                    for (int o = 0; o < n_max; o += 8) {
                        int idx[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) idx[k] = n_mine > 0 ? __ldg(A.col_src + e0 + min(o + k, last)) : 0;
                        float4 v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (o + k < n_mine) v[k] = ld_row4(A.xin, idx[k], q);
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) f4add(sum, v[k]);
                    }
                }
                if (!TGNN_TIMED(w1, mbar_wait(bar_ce + 8 * mw, (uint32_t)((use & 1) ^ 1)))) { timeout_flag = 1; ok = false; break; }   // its previous chunk is in registers
                *reinterpret_cast<float4*>(S + (4 * (qd & 3) + a) * XS + 4 * q) = sum;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_cf + 8 * mw);
            }
            if (!ok) break;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_we + 8 * b);
        }
    } else {
        // ===================== node MLP: chunk c of the CTA's it-th tile goes to warp (4 it + c) % 11 =====================
        const GinW<HMLP> Wt(wsm);
        double s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] = 0.0; s2[j] = 0.0; }
        const float* S = reinterpret_cast<const float*>(smem + OFF_S) + warp * (CH * XS);
        int use = 0;
        for (int ci = warp;; ci += MLP_WARPS, ++use) {
            const int it = ci / CHUNKS_PER_TILE, c = ci % CHUNKS_PER_TILE;
            const long long tile = (long long)blockIdx.x + (long long)it * gridDim.x;
            if (tile >= n_tiles) break;
            if (!TGNN_TIMED(w0, mbar_wait(bar_cf + 8 * warp, (uint32_t)(use & 1)))) { timeout_flag = 1; break; }
            float a1[4][4];
            gin_load_a1(S, lane, a1);              // into registers, then the buffer goes back to the gather warps
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ce + 8 * warp);
            const int node0 = (int)tile * GW_T + c * CH;
            if (node0 < A.n_own) gin_mlp_chunk<HMLP>(a1, Wt, node0, A.n_own, A.out, s1, s2, lane, A.mask);
        }
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
        }
        if (A.part && g == 0) {
            const size_t row = (size_t)blockIdx.x * MLP_WARPS + warp;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                A.part[row * 64 + 8 * t + j] = s1[j];
                A.part[row * 64 + 32 + 8 * t + j] = s2[j];
            }
        }
    }
    if (A.dbg && blockIdx.x == 0 && lane == 0) {
        long long* d = A.dbg + warp * 4;
        d[0] = clock64() - t_start; d[1] = w0; d[2] = w1; d[3] = w2;
    }
    __syncthreads();
    if (timeout_flag && tid == 0 && A.err) { *reinterpret_cast<volatile int*>(A.err) = TGNN_DEVERR_PIPELINE; __threadfence_system(); }
}

// ---- window builder: one CTA per 64-row tile sorts the tile's source rows (plus its own rows, for the self term) in
// ---- shared memory, cuts them into contiguous runs (gaps of <= GW_GAP rows are loaded rather than split) and rewrites
// ---- every edge's source as its offset inside the window -------------------------------------------------------------
constexpr int GW_ITEMS = GW_CAP / 256;

__global__ void __launch_bounds__(256)
k_gw_build(const int* __restrict__ col_ptr, const int* __restrict__ col_src, int n_own, int* __restrict__ meta,
           int* __restrict__ seg, uint16_t* __restrict__ loc, int* __restrict__ n_direct) {
    using Sort = cub::BlockRadixSort<int, 256, GW_ITEMS>;
    using Scan = cub::BlockScan<int, 256>;
    __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
    __shared__ int srt[GW_CAP];
    __shared__ int seg_first[GW_MAXSEG], seg_lbase[GW_MAXSEG];
    __shared__ int s_nseg, s_rows;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int node0 = tile * GW_T, node1 = min(node0 + GW_T, n_own);
    const int e0 = col_ptr[node0], e1 = col_ptr[node1];
    const int ne = e1 - e0, n_items = ne + (node1 - node0);
    const int off = (e0 + 8 * tile) & ~7;
    int* m = meta + 8 * tile;
    if (n_items > GW_CAP) {
        if (tid == 0) { m[0] = 0; m[1] = 0; m[2] = off; m[3] = 0; m[4] = ne; atomicAdd(n_direct, 1); }
        return;
    }
    int items[GW_ITEMS];
#pragma unroll
    for (int k = 0; k < GW_ITEMS; ++k) {
        const int i = tid * GW_ITEMS + k;
        items[k] = i < ne ? col_src[e0 + i] : (i < n_items ? node0 + (i - ne) : 0x7fffffff);
    }
    Sort(tmp.sort).Sort(items);
#pragma unroll
    for (int k = 0; k < GW_ITEMS; ++k) srt[tid * GW_ITEMS + k] = items[k];
    __syncthreads();
    // run starts and the rows skipped in front of each run ("jump"): local(row) = row - (sum of jumps up to it)
    int jump[GW_ITEMS], start[GW_ITEMS], jsum = 0, ssum = 0;
#pragma unroll
    for (int k = 0; k < GW_ITEMS; ++k) {
        const int i = tid * GW_ITEMS + k;
        const int v = items[k], prev = i > 0 ? srt[i - 1] : 0;
        const bool valid = i < n_items;
        const bool st = valid && (i == 0 || v - prev - 1 > GW_GAP);
        start[k] = st ? 1 : 0;
        jump[k] = st ? (i == 0 ? v : v - prev - 1) : 0;
        jsum += jump[k]; ssum += start[k];
    }
    int jpre, spre;
    Scan(tmp.scan).ExclusiveSum(jsum, jpre);
    __syncthreads();
    Scan(tmp.scan).ExclusiveSum(ssum, spre);
    if (tid == 0) { s_nseg = 0; s_rows = 0; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GW_ITEMS; ++k) {
        const int i = tid * GW_ITEMS + k;
        jpre += jump[k]; spre += start[k];
        if (start[k] && spre <= GW_MAXSEG) { seg_first[spre - 1] = items[k]; seg_lbase[spre - 1] = items[k] - jpre; }
        if (i == n_items - 1) { s_nseg = spre; s_rows = items[k] - jpre + 1; }
    }
    __syncthreads();
    const int nseg = s_nseg, rows = s_rows;
    if (nseg > GW_MAXSEG || rows > GW_WMAX) {
        if (tid == 0) { m[0] = 0; m[1] = 0; m[2] = off; m[3] = 0; m[4] = ne; atomicAdd(n_direct, 1); }
        return;
    }
    auto local_of = [&](int row) {
        int lo = 0, hi = nseg - 1;                       // last run whose first row is <= row
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (seg_first[mid] <= row) lo = mid; else hi = mid - 1; }
        return seg_lbase[lo] + (row - seg_first[lo]);
    };
    for (int i = tid; i < ne; i += 256) loc[off + i] = (uint16_t)local_of(col_src[e0 + i]);
    if (tid < nseg) { seg[((size_t)tile * GW_MAXSEG + tid) * 2] = seg_first[tid]; seg[((size_t)tile * GW_MAXSEG + tid) * 2 + 1] = seg_lbase[tid]; }
    if (tid == 0) { m[0] = nseg; m[1] = rows; m[2] = off; m[3] = local_of(node0); m[4] = ne; }
}

}  // namespace

// Builds the windows of the collision CSR already in g (col_ptr / col_src).  Returns the number of "direct" tiles.
int build_gin_windows(Graph& g, Scratch& sc, cudaStream_t st) {
    g.gw_tiles = (int)((g.n_own + GW_T - 1) / GW_T);
    g.gw_meta.reserve((size_t)g.gw_tiles * 8 * sizeof(int));
    g.gw_seg.reserve((size_t)g.gw_tiles * GW_MAXSEG * 2 * sizeof(int));
    g.gw_loc.reserve(((size_t)g.e_col + 8 * (size_t)g.gw_tiles + 64) * sizeof(uint16_t));
    int* nd = sc.get<int>(1);
    TGNN_CUDA(cudaMemsetAsync(nd, 0, sizeof(int), st));
    k_gw_build<<<g.gw_tiles, 256, 0, st>>>(g.col_ptr.as<int>(), g.col_src.as<int>(), (int)g.n_own, g.gw_meta.as<int>(),
                                           g.gw_seg.as<int>(), g.gw_loc.as<uint16_t>(), nd);
    TGNN_CUDA(cudaGetLastError());
    int n_direct = 0;
    TGNN_CUDA(cudaMemcpyAsync(&n_direct, nd, sizeof(int), cudaMemcpyDeviceToHost, st));
    TGNN_CUDA(cudaStreamSynchronize(st));
    return n_direct;
}

int gin_w_blocks(int gw_tiles, int sm_count) { return gw_tiles < sm_count ? (gw_tiles < 1 ? 1 : gw_tiles) : sm_count; }

void launch_gin_w(const GinArgs& a, int sm_count, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([&] {
        TGNN_CUDA(cudaFuncSetAttribute(k_gin_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gw_smem_bytes<true>()));
    });
    // only the fp16-MLP variant exists: the 3xTF32 weight tables (41 KB) do not fit next to two window buffers, and
    // layers whose GIN weights are outside the fp16 range stay on k_gin (the caller checks a.hmlp)
    TGNN_CHECK(a.hmlp, "internal: k_gin_w needs the fp16 MLP tables");
    const int blocks = gin_w_blocks(a.gw_tiles, sm_count);
    k_gin_w<true><<<blocks, GW_THREADS, gw_smem_bytes<true>(), st>>>(a);
    TGNN_CUDA(cudaGetLastError());
}

}  // namespace tgnn
