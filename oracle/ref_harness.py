"""Run the reference's OWN ``graph_networks`` code (test infrastructure).

Only usable where ``/root/reference`` exists (the build container); it is how
``tests/golden/make_golden.py`` produces the vectors that pin
``oracle/tilingnn_oracle.py``.  Nothing here runs on the GPU box.

``torch_geometric`` (README.md:10-11 of the reference: "tested with v1.3.2") is
not installable offline, so the two classes the reference imports
(/root/reference/graph_networks/layers/edge_conv.py:3 and coll_conv.py:3) are
provided as pure-torch stand-ins that follow PyG 1.3.x's published algorithm
and attribute names (``nn``, ``root``, ``bias``, ``eps``), which makes the
shipped checkpoints load with ``strict=True``.  ``inputs.config`` is stubbed
because importing the real one needs shapely and cwd = the reference root
(/root/reference/inputs/config.py:26); only ``environment.tile_count`` is read
(/root/reference/graph_networks/networks/TilinGNN.py:19).
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


class NNConv(nn.Module):
    """PyG 1.3.x ``NNConv``: message = x_j @ nn(e).view(in,out); aggr; + x@root + bias."""

    def __init__(self, in_channels, out_channels, nn_module, aggr="add", root_weight=True, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.aggr = in_channels, out_channels, aggr
        self.nn = nn_module
        self.root = nn.Parameter(torch.empty(in_channels, out_channels)) if root_weight else None
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        bound = 1.0 / (in_channels ** 0.5)
        for t in (self.root, self.bias):
            if t is not None:
                nn.init.uniform_(t, -bound, bound)

    def forward(self, x, edge_index, edge_attr):
        src, dst = edge_index[0], edge_index[1]
        w = self.nn(edge_attr).view(-1, self.in_channels, self.out_channels)
        msg = torch.matmul(x[src].unsqueeze(1), w).squeeze(1)
        out = torch.zeros(x.shape[0], self.out_channels, dtype=x.dtype, device=x.device)
        out.index_add_(0, dst, msg)
        if self.aggr == "mean":
            cnt = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
            cnt.index_add_(0, dst, torch.ones_like(dst, dtype=x.dtype))
            out = out / cnt.clamp(min=1).unsqueeze(1)
        if self.root is not None:
            out = out + torch.mm(x, self.root)
        if self.bias is not None:
            out = out + self.bias
        return out


class GINConv(nn.Module):
    """PyG 1.3.x ``GINConv``: nn((1+eps)*x + sum_{j != i} x_j), eps a buffer."""

    def __init__(self, nn, eps=0.0, train_eps=False):
        super().__init__()
        self.nn = nn
        self.register_buffer("eps", torch.Tensor([eps]))

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        keep = src != dst
        agg = torch.zeros_like(x)
        agg.index_add_(0, dst[keep], x[src[keep]])
        return self.nn((1 + self.eps) * x + agg)


def install_shims(tile_count=2):
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_conv = types.ModuleType("torch_geometric.nn.conv")
    tg_nnconv = types.ModuleType("torch_geometric.nn.conv.nn_conv")
    tg_nn.GINConv, tg_nn.NNConv = GINConv, NNConv
    tg_nnconv.NNConv = NNConv
    tg.nn, tg_nn.conv, tg_conv.nn_conv = tg_nn, tg_conv, tg_nnconv
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn,
                        "torch_geometric.nn.conv": tg_conv,
                        "torch_geometric.nn.conv.nn_conv": tg_nnconv})
    inputs = types.ModuleType("inputs")
    cfg = types.ModuleType("inputs.config")
    cfg.environment = types.SimpleNamespace(tile_count=tile_count)
    inputs.config = cfg
    sys.modules.update({"inputs": inputs, "inputs.config": cfg})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_network(d_x, d_e, depth, width=32):
    """Construct the reference ``TilinGNN`` (unmodified source)."""
    install_shims(tile_count=d_x - 1)
    from graph_networks.networks.TilinGNN import TilinGNN
    return TilinGNN(adj_edge_features_dim=d_e, network_depth=depth, network_width=width,
                    node_features_dim=d_x)


def run_reference(state_dict, x, adj_e_index, adj_e_features, col_e_idx, *, depth,
                  bn_mode="train", dtype=torch.float64, width=32):
    """One forward of the reference module in ``dtype`` on CPU; returns scores [N,1]."""
    d_x, d_e = x.shape[1], adj_e_features.shape[1]
    net = reference_network(d_x, d_e, depth, width)
    net.load_state_dict(state_dict, strict=True)
    net = net.to(dtype)
    net.train() if bn_mode == "train" else net.eval()
    with torch.no_grad():
        probs, *_ = net(x=x.to(dtype), adj_e_index=adj_e_index.long(),
                        adj_e_features=adj_e_features.to(dtype), col_e_idx=col_e_idx.long(),
                        col_e_features=None)
    return probs
