"""``TilinGNN`` -- drop-in for /root/reference/graph_networks/networks/TilinGNN.py.

Same constructor, same ``forward(x, adj_e_index, adj_e_features, col_e_idx, col_e_features=None)``
signature and return value ``(scores[N,1], adj_e_features)``, same ``state_dict`` keys (all 664 of
the shipped checkpoints, including the aliased ``nnConv.nn.mlp.*`` entries and the BatchNorm
buffers), ``.to()``, ``.train()/.eval()`` and ``copy.deepcopy`` -- but the arithmetic runs in the
hand-written sm_100a kernels of ``libtgnn.so`` through its C ABI (include/tgnn.h).  The torch
modules below only HOLD parameters under the reference's names; they never compute.

There is no CPU or PyTorch fallback: calling ``forward`` with CPU tensors, or without the built
library, raises.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


# --------------------------------------------------------------------------------------------- #
# parameter holders that reproduce the reference's module tree (names only)                      #
# --------------------------------------------------------------------------------------------- #
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: the arithmetic lives in libtgnn.so, call TilinGNN.forward")


class Linear_trans(_Holder):
    """Names of graph_networks/layers/util.py:21-29 (``linear``, ``batch_norm``)."""

    def __init__(self, in_dim, out_dim, batch_norm=True):
        super().__init__()
        self.linear = nn.Linear(in_dim, out_dim)
        if batch_norm:
            self.batch_norm = nn.BatchNorm1d(out_dim)


class MLP(_Holder):
    """Names of graph_networks/layers/util.py:4-13 (``mlp`` = Sequential of Linear_trans)."""

    def __init__(self, in_dim, out_dim, hidden_layer_dims, batch_norm=True):
        super().__init__()
        dims = [in_dim] + list(hidden_layer_dims) + [out_dim]
        self.mlp = nn.Sequential(*[Linear_trans(dims[i], dims[i + 1], batch_norm) for i in range(len(dims) - 1)])


class _NNConvParams(_Holder):
    """PyG ``NNConv`` attribute names: ``root`` [in,out], ``bias`` [out], ``nn``."""

    def __init__(self, in_channels, out_channels, nn_module):
        super().__init__()
        bound = 1.0 / (in_channels ** 0.5)
        self.root = nn.Parameter(torch.empty(in_channels, out_channels).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        self.nn = nn_module


class _GINConvParams(_Holder):
    """PyG ``GINConv`` attribute names: buffer ``eps`` [1], ``nn``."""

    def __init__(self, nn_module, eps=0.0):
        super().__init__()
        self.register_buffer("eps", torch.tensor([eps], dtype=torch.float32))
        self.nn = nn_module


class GraphConv(_Holder):
    """Names of graph_networks/layers/edge_conv.py:7-22."""

    def __init__(self, edge_feature_dim, node_feature_in_dim, node_feature_out_dim, hidden_dims=(32, 64)):
        super().__init__()
        self.mlp = MLP(edge_feature_dim, node_feature_in_dim * node_feature_out_dim, hidden_dims, batch_norm=False)
        self.nnConv = _NNConvParams(node_feature_in_dim, node_feature_out_dim, self.mlp)
        self.batch_norm = nn.BatchNorm1d(node_feature_out_dim)


class CollConv(_Holder):
    """Names of graph_networks/layers/coll_conv.py:7-22."""

    def __init__(self, node_feature_in_dim, node_feature_out_dim, hidden_dims=(32, 64)):
        super().__init__()
        self.ginConv = _GINConvParams(MLP(node_feature_in_dim, node_feature_out_dim, hidden_dims, batch_norm=False))
        self.batch_norm = nn.BatchNorm1d(node_feature_out_dim)


class _Native:
    """Owner of the C handle.  Never copied: a deep copy of the module starts with a fresh one."""

    def __init__(self):
        self.h = None
        self.device_index = None
        self.params_synced = False
        self.graph_key = None
        self.keepalive = None
        self.bn_mode = None
        self.shard = None            # (rank, world) once tgnn_shard_init has run

    def __deepcopy__(self, memo):
        return _Native()

    def __getstate__(self):          # pickling a module must not carry a raw pointer
        return {}

    def __setstate__(self, state):
        self.__init__()

    def close(self):
        if self.h is not None:
            try:
                _lib.load().tgnn_destroy(self.h)
            except Exception:
                pass
            self.h = None

    def __del__(self):
        self.close()


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() > 0 else 0)


class TilinGNN(nn.Module):
    """See the module docstring.  Constructor mirrors TilinGNN.py:14-20 of the reference
    (``node_features_dim`` has no config-derived default here: pass ``tile_count + 1``)."""

    def __init__(self, adj_edge_features_dim, network_depth, network_width, output_dim=1, node_features_dim=None):
        super().__init__()
        if node_features_dim is None:
            raise TypeError("node_features_dim is required (the reference derives it from inputs.config: "
                            "environment.tile_count + 1)")
        if network_width != 32:
            raise ValueError("tilingnn_b200 implements network_width = 32 (inputs/config.py:38 of the reference)")
        if output_dim != 1:
            raise ValueError("tilingnn_b200 implements output_dim = 1 (num_prob_maps = 1 everywhere in the reference)")
        self.network_depth = network_depth
        self.network_width = network_width
        self.residual_skip_num = 2
        self.adj_edge_features_dim = adj_edge_features_dim
        self.node_features_dim = node_features_dim
        W, L = network_width, network_depth
        self.init_node_feature_trans = MLP(node_features_dim, W, [W], batch_norm=True)
        self.brch_1_graph_conv_layers = nn.ModuleList([GraphConv(adj_edge_features_dim, W, W) for _ in range(L)])
        self.brch_2_coll_conv_layers = nn.ModuleList([CollConv(W, W) for _ in range(L)])
        self.final_mlp = nn.Sequential(MLP(W * (L + 1), W, [256, 128, 64], batch_norm=True),
                                       Linear_trans(W, output_dim, batch_norm=False))
        self._native = _Native()

    # ---- parameter / device bookkeeping --------------------------------------------------------
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._native.params_synced = False
        return out

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._native.params_synced = False
        return out

    def mark_parameters_dirty(self):
        """Call after modifying parameters in place (e.g. an optimiser step)."""
        self._native.params_synced = False

    def _device(self):
        return self.final_mlp[1].linear.weight.device

    def _ensure_handle(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("tilingnn_b200.TilinGNN runs on CUDA (sm_100a) only; move the module with "
                               ".to('cuda') -- there is no CPU fallback")
        nat = self._native
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if nat.h is not None and nat.device_index == idx:
            return nat
        nat.close()
        lib = _lib.load()
        cfg = _lib.tgnn_cfg(self.node_features_dim, self.adj_edge_features_dim, self.network_width, self.network_depth,
                            _lib.TGNN_BN_TRAIN if self.training else _lib.TGNN_BN_EVAL, idx)
        h = C.c_void_p()
        _lib.check(None, lib.tgnn_create(C.byref(cfg), C.byref(h)), "tgnn_create")
        nat.h, nat.device_index = h, idx
        nat.params_synced, nat.graph_key, nat.bn_mode, nat.shard = False, None, None, None
        return nat

    def _sync(self):
        nat = self._ensure_handle()
        lib = _lib.load()
        if not nat.params_synced:
            for key, t in self.state_dict().items():
                if not t.is_floating_point():
                    continue
                t32 = t.detach().to(torch.float32).contiguous()
                shape = (C.c_int64 * max(1, t32.dim()))(*t32.shape)
                _lib.check(nat.h, lib.tgnn_set_param(nat.h, key.encode(), _ptr(t32), shape, t32.dim()),
                           "tgnn_set_param")
            nat.params_synced = True
        mode = _lib.TGNN_BN_TRAIN if self.training else _lib.TGNN_BN_EVAL
        if nat.bn_mode != mode:
            _lib.check(nat.h, lib.tgnn_set_bn_mode(nat.h, mode), "tgnn_set_bn_mode")
            nat.bn_mode = mode
        return nat

    # ---- graph ---------------------------------------------------------------------------------
    @staticmethod
    def _edge_rows(index):
        if index is None or index.numel() == 0:
            return None, None, 0
        if index.dim() != 2 or index.shape[0] != 2:
            raise ValueError(f"edge index must have shape [2, E], got {tuple(index.shape)}")
        index = index.to(torch.int64)
        src, dst = index[0].contiguous(), index[1].contiguous()
        return src, dst, src.numel()

    def set_graph(self, num_nodes, adj_e_index, adj_e_features, col_e_idx):
        """Build the device-side graph structures once; ``score(x)`` then reuses them.
        (``forward`` calls this itself, keyed on the identity of the index tensors.)"""
        nat = self._sync()
        lib = _lib.load()
        dev = self._device()
        for t in (adj_e_index, adj_e_features, col_e_idx):
            if t is not None and t.numel() > 0 and t.device != dev:
                raise RuntimeError(f"graph tensors must be on {dev}, got {t.device} (no CPU fallback)")
        a_src, a_dst, e_a = self._edge_rows(adj_e_index)
        c_src, c_dst, e_c = self._edge_rows(col_e_idx)
        feat = None
        if e_a > 0:
            feat = adj_e_features.to(torch.float32).contiguous()
            if feat.dim() != 2 or feat.shape[0] != e_a or feat.shape[1] != self.adj_edge_features_dim:
                raise ValueError(f"adj_e_features must be [{e_a}, {self.adj_edge_features_dim}], got {tuple(feat.shape)}")
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.tgnn_set_graph(nat.h, int(num_nodes), e_a, _ptr(a_src), _ptr(a_dst), _ptr(feat),
                                    e_c, _ptr(c_src), _ptr(c_dst), C.c_void_p(st))
        nat.graph_key = None
        _lib.check(nat.h, rc, "tgnn_set_graph")
        nat.num_nodes = int(num_nodes)
        nat.graph_serial = getattr(nat, "graph_serial", 0) + 1      # lets callers notice that their resident graph was replaced

    def score(self, x, out=None):
        """One scoring pass on the resident graph: x [N, d_x] (CUDA) -> scores [N] fp32."""
        nat = self._sync()
        lib = _lib.load()
        dev = self._device()
        if x.device != dev:
            raise RuntimeError(f"x must be on {dev}, got {x.device} (no CPU fallback)")
        x32 = x.detach().to(torch.float32).contiguous()
        if x32.dim() != 2 or x32.shape[1] != self.node_features_dim:
            raise ValueError(f"x must be [N, {self.node_features_dim}], got {tuple(x32.shape)}")
        n = x32.shape[0]
        if getattr(nat, "num_nodes", None) != n:
            raise RuntimeError("score(): x has a different node count than the resident graph")
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.tgnn_forward(nat.h, _ptr(x32), _ptr(out), C.c_void_p(st))
        _lib.check(nat.h, rc, "tgnn_forward")
        return out

    def set_node_mask(self, keep):
        """Sub-layout on the RESIDENT graph (tgnn_set_node_mask; reference: BrickLayout.compute_sub_layout,
        tiling/brick_layout.py:248-286, without the re-indexing): ``keep`` is a bool / uint8 vector [N] (torch, CPU or
        CUDA, or numpy), ``None`` restores the full graph.  The next ``score(x)`` calls score the sub-graph induced by the
        kept nodes; masked nodes get score 0.  Returns ``(kept nodes, adjacency edges, collision edges)`` of the sub-graph."""
        nat = self._sync()
        lib = _lib.load()
        dev = self._device()
        counts = (C.c_int64 * 3)()
        st = torch.cuda.current_stream(dev).cuda_stream
        if keep is None:
            ptr, hold = C.c_void_p(0), None
        else:
            hold = torch.as_tensor(keep)
            if hold.dtype != torch.uint8:
                hold = hold.to(torch.uint8)
            hold = hold.contiguous()
            if hold.numel() != getattr(nat, "num_nodes", None):
                raise ValueError("set_node_mask: keep must have one entry per node of the resident graph")
            ptr = _ptr(hold)
        with torch.cuda.device(dev):
            rc = lib.tgnn_set_node_mask(nat.h, ptr, counts, C.c_void_p(st))      # synchronises (the counts come back)
        _lib.check(nat.h, rc, "tgnn_set_node_mask")
        return int(counts[0]), int(counts[1]), int(counts[2])

    def check_errors(self, synchronize=True):
        """Raise if a kernel of an earlier forward reported a device-side failure (tcgen05 pipeline or peer-exchange
        timeout).  ``tgnn_forward`` checks by itself for small graphs and at the next call for large ones; call this
        after the last forward of a batch of large graphs."""
        nat = self._ensure_handle()
        dev = self._device()
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = _lib.load().tgnn_check_error(nat.h, C.c_void_p(st), int(bool(synchronize)))
        _lib.check(nat.h, rc, "tgnn_check_error")

    # ---- multi-GPU: node-range shards --------------------------------------------
    def shard_init(self, group=None):
        """Join this module's handle to an NCCL communicator of its own (one process per GPU).
        ``torch.distributed`` must be initialised; it only carries the 128-byte unique id."""
        import torch.distributed as dist
        nat = self._sync()
        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = self._device()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            _lib.check(None, lib.tgnn_nccl_unique_id(buf), "tgnn_nccl_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        carrier = uid.to(dev) if dist.get_backend(group) == "nccl" else uid
        dist.broadcast(carrier, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(carrier.cpu().tolist())
        with torch.cuda.device(dev):
            rc = lib.tgnn_shard_init(nat.h, C.c_char_p(raw), rank, world)
        _lib.check(nat.h, rc, "tgnn_shard_init")
        nat.shard = (rank, world)

    def set_graph_shard(self, plan, adj_e_features):
        """``plan``: tilingnn_b200.shard.ShardPlan for this rank; ``adj_e_features`` [E_a, d_e] of the
        rank's own adjacency edges (same order as the plan's edge arrays)."""
        nat = self._sync()
        if nat.shard is None:
            raise RuntimeError("set_graph_shard: call shard_init() first")
        lib = _lib.load()
        dev = self._device()
        mv = lambda t: t.to(dev).to(torch.int64).contiguous()
        a_src, a_dst = mv(plan.adj_src_local), mv(plan.adj_dst_local)
        c_src, c_dst = mv(plan.col_src_local), mv(plan.col_dst_local)
        send = mv(plan.send_rows)
        feat = adj_e_features.to(dev).to(torch.float32).contiguous() if a_src.numel() else None
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.tgnn_set_graph_shard(nat.h, plan.n_own, plan.n_global, plan.halo_slot, send.numel(), _ptr(send),
                                          a_src.numel(), _ptr(a_src), _ptr(a_dst), _ptr(feat),
                                          c_src.numel(), _ptr(c_src), _ptr(c_dst), C.c_void_p(st))
        nat.graph_key = None
        _lib.check(nat.h, rc, "tgnn_set_graph_shard")
        if getattr(plan, "send_mask", None) is not None and send.numel() > 0:
            sm = plan.send_mask.to(dev).to(torch.uint8).contiguous()
            with torch.cuda.device(dev):
                rc = lib.tgnn_set_halo_peers(nat.h, _ptr(sm), sm.numel(), C.c_void_p(st))
            _lib.check(nat.h, rc, "tgnn_set_halo_peers")
        nat.num_nodes = plan.n_own

    # ---- the reference signature ---------------------------------------------------------------
    def forward(self, x, adj_e_index, adj_e_features, col_e_idx, col_e_features=None):
        nat = self._ensure_handle()
        key = tuple((t.data_ptr(), tuple(t.shape), t._version) if t is not None else None
                    for t in (adj_e_index, adj_e_features, col_e_idx)) + (int(x.shape[0]),)
        if nat.graph_key != key:
            self.set_graph(x.shape[0], adj_e_index, adj_e_features, col_e_idx)
            nat.graph_key = key
            nat.keepalive = (adj_e_index, adj_e_features, col_e_idx)   # data_ptr identity stays valid
        scores = self.score(x)
        return scores.view(-1, 1), adj_e_features

    # ---- introspection -------------------------------------------------------------------------
    def info(self):
        nat = self._ensure_handle()
        inf = _lib.tgnn_info()
        _lib.check(nat.h, _lib.load().tgnn_get_info(nat.h, C.byref(inf)), "tgnn_get_info")
        return {n: getattr(inf, n) for n, _ in _lib.tgnn_info._fields_}

    def set_profiling(self, enabled):
        nat = self._ensure_handle()
        _lib.check(nat.h, _lib.load().tgnn_set_profiling(nat.h, int(bool(enabled))), "tgnn_set_profiling")

    def profile(self):
        nat = self._ensure_handle()
        out = {}
        for fam in ("init", "conv", "gin", "bnfin", "combine", "final", "score", "halo"):
            ms, n = C.c_float(), C.c_int32()
            _lib.check(nat.h, _lib.load().tgnn_get_profile(nat.h, fam.encode(), C.byref(ms), C.byref(n)), "tgnn_get_profile")
            out[fam] = (ms.value, n.value)
        return out

    def debug_set_stop_layer(self, layer):
        nat = self._ensure_handle()
        _lib.check(nat.h, _lib.load().tgnn_debug_set_stop_layer(nat.h, int(layer)), "tgnn_debug_set_stop_layer")

    def debug_read(self, name):
        nat = self._ensure_handle()
        dev = self._device()
        out = torch.empty(nat.num_nodes, 32, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(nat.h, _lib.load().tgnn_debug_read(nat.h, name.encode(), _ptr(out), C.c_void_p(st)), "tgnn_debug_read")
        torch.cuda.synchronize(dev)
        return out

    def debug_graph(self):
        """Built graph structures as CPU tensors (tests)."""
        nat = self._ensure_handle()
        inf = self.info()
        n_tiles = (inf["n_own"] + inf["tile_rows"] - 1) // inf["tile_rows"]
        n_chunks = inf["adj_slots"] // 16
        t = dict(cptr=torch.zeros(n_tiles + 1, dtype=torch.int32), ctype=torch.zeros(n_chunks, dtype=torch.int32),
                 csrc=torch.zeros(n_chunks * 16, dtype=torch.int32), cdst=torch.zeros(n_chunks * 16, dtype=torch.uint8),
                 inv_deg=torch.zeros(inf["n_own"], dtype=torch.float32),
                 col_ptr=torch.zeros(inf["n_own"] + 1, dtype=torch.int32),
                 col_src=torch.zeros(inf["e_col"], dtype=torch.int32),
                 type_rows=torch.zeros(inf["n_edge_types"], self.adj_edge_features_dim, dtype=torch.float32))
        order = ("cptr", "ctype", "csrc", "cdst", "inv_deg", "col_ptr", "col_src", "type_rows")
        rc = _lib.load().tgnn_debug_graph(nat.h, *[_ptr(t[k]) for k in order], C.c_void_p(0))
        _lib.check(nat.h, rc, "tgnn_debug_graph")
        return t

    def debug_graph_t(self):
        """The edge-block format of the tcgen05 adjacency kernel as CPU tensors (tests)."""
        nat = self._ensure_handle()
        inf = self.info()
        if not inf["t_rows"]:
            raise RuntimeError("debug_graph_t: the edge-block format was not built for this graph")
        n_tiles = (inf["n_own"] + inf["t_rows"] - 1) // inf["t_rows"]
        t = dict(bptr=torch.zeros(n_tiles + 1, dtype=torch.int32), btype=torch.zeros(inf["t_blocks"], dtype=torch.int32),
                 tsrc=torch.zeros(inf["t_blocks"] * 128, dtype=torch.int32), tdst=torch.zeros(inf["t_blocks"] * 128, dtype=torch.int16))
        rc = _lib.load().tgnn_debug_graph_t(nat.h, *[_ptr(t[k]) for k in ("bptr", "btype", "tsrc", "tdst")], C.c_void_p(0))
        _lib.check(nat.h, rc, "tgnn_debug_graph_t")
        t["tdst"] = t["tdst"].to(torch.int32) & 0xFFFF
        return t

    def debug_role_cycles(self):
        """TGNN_ROLE_DBG=1: {kernel: [[cycles, wait0, wait1, wait2] per warp]} of CTA 0 in the last forward."""
        nat = self._ensure_handle()
        buf = (C.c_int64 * 256)()
        _lib.check(nat.h, _lib.load().tgnn_debug_role_cycles(nat.h, buf), "tgnn_debug_role_cycles")
        v = list(buf)
        return {"k_conv_t|z": [v[4 * w: 4 * w + 4] for w in range(16)], "k_gin_w": [v[128 + 4 * w: 128 + 4 * w + 4] for w in range(16)]}
