"""CPU oracle for the TilinGNN per-node scoring forward pass.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tilingnn_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker / the CPU baseline, never as the product path.

What it is: a functional restatement, on CPU torch tensors (fp64 = gold,
fp32 = "the reference as shipped"), of the arithmetic of

  * ``TilinGNN.forward``      /root/reference/graph_networks/networks/TilinGNN.py:51-78
  * ``GraphConv.forward``     /root/reference/graph_networks/layers/edge_conv.py:24-30
  * ``CollConv.forward``      /root/reference/graph_networks/layers/coll_conv.py:24-30
  * ``MLP`` / ``Linear_trans``/root/reference/graph_networks/layers/util.py:4-37
  * ``nn.BatchNorm1d`` in train mode (mode set by
    /root/reference/solver/ml_solver/ml_solver.py:129-131) or eval mode.

Third-party arithmetic that is NOT under /root/reference (no lock file; the
versions are stated in prose at /root/reference/README.md:10-11):
``torch_geometric`` 1.3.2 ``NNConv(aggr="mean")`` and ``GINConv``.  Their
published algorithms are restated in ``nnconv_mean`` and ``ginconv`` below
(call sites: edge_conv.py:18,25 and coll_conv.py:18,25).

Parity pin: the reference has no golden vectors for this path.
The oracle is pinned against outputs of the reference's OWN ``graph_networks``
modules, imported unmodified in the build container (PyG classes shimmed, see
``oracle/ref_harness.py``) on the shipped checkpoints and graphs; the vectors
are committed under ``tests/golden/`` together with the generating script
``tests/golden/make_golden.py`` and checked by ``tests/test_oracle_golden.py``.

Parameters are passed as a ``dict`` keyed by the reference's ``state_dict``
names, so the same checkpoint feeds the reference, the
oracle and the CUDA path.
"""
from __future__ import annotations

import torch

_BN_CAPTURE = None      # set by forward(_capture_bn=True): {bn prefix: (mean, biased var)}
LEAKY_SLOPE = 0.01      # torch.nn.LeakyReLU() default, TilinGNN.py:31,46
BN_EPS = 1e-5           # torch.nn.BatchNorm1d default


def _lin(h, p, prefix):
    # nn.Linear: y = x W^T + b  (util.py:24,32)
    return h @ p[prefix + ".weight"].t() + p[prefix + ".bias"]


def _leaky(h):
    return torch.where(h >= 0, h, h * LEAKY_SLOPE)


def batch_norm(h, p, prefix, bn_mode):
    """BatchNorm1d over the rows of THIS call (train) or running stats (eval)."""
    if bn_mode == "train":
        mu = h.mean(dim=0)
        var = ((h - mu) ** 2).mean(dim=0)          # biased, as F.batch_norm uses for normalisation
        if _BN_CAPTURE is not None:
            _BN_CAPTURE[prefix] = (mu.clone(), var.clone())
    else:
        mu = p[prefix + ".running_mean"]
        var = p[prefix + ".running_var"]
    return (h - mu) / torch.sqrt(var + BN_EPS) * p[prefix + ".weight"] + p[prefix + ".bias"]


def linear_trans(h, p, prefix, act, bn, bn_mode):
    """``Linear_trans.forward`` (util.py:31-37): linear -> activation -> BN."""
    h = _lin(h, p, prefix + ".linear")
    if act == "leaky":
        h = _leaky(h)
    elif act == "sigmoid":
        h = torch.sigmoid(h)
    if bn:
        h = batch_norm(h, p, prefix + ".batch_norm", bn_mode)
    return h


def mlp(h, p, prefix, n_layers, act, bn, bn_mode):
    """``MLP.forward`` (util.py:15-17): the same activation after EVERY layer."""
    for k in range(n_layers):
        h = linear_trans(h, p, f"{prefix}.mlp.{k}", act, bn, bn_mode)
    return h


def nnconv_mean(x, edge_index, edge_attr, p, prefix, edge_chunk=1 << 16):
    """PyG 1.3.x ``NNConv(in, out, nn, aggr="mean")``.

    flow = source_to_target: x_j = x[edge_index[0]], aggregated at edge_index[1];
    weight = nn(edge_attr).view(-1, in, out); msg = x_j[:,None,:] @ weight;
    scatter_mean with dim_size = N (rows without in-edges stay 0);
    update = aggr + x @ root + bias.
    """
    n, f_in = x.shape
    f_out = p[prefix + ".nnConv.bias"].numel()
    src, dst = edge_index[0], edge_index[1]
    agg = torch.zeros(n, f_out, dtype=x.dtype)
    cnt = torch.zeros(n, dtype=x.dtype)
    e_total = src.numel()
    for lo in range(0, e_total, edge_chunk):       # chunked only to bound the [E, in*out] tensor
        hi = min(e_total, lo + edge_chunk)
        w = mlp(edge_attr[lo:hi], p, prefix + ".mlp", 3, "sigmoid", False, "train")
        w = w.view(-1, f_in, f_out)
        msg = torch.matmul(x[src[lo:hi]].unsqueeze(1), w).squeeze(1)
        agg.index_add_(0, dst[lo:hi], msg)
        cnt.index_add_(0, dst[lo:hi], torch.ones(hi - lo, dtype=x.dtype))
    agg = agg / cnt.clamp(min=1).unsqueeze(1)
    return agg + x @ p[prefix + ".nnConv.root"] + p[prefix + ".nnConv.bias"]


def ginconv(x, edge_index, p, prefix):
    """PyG 1.3.x ``GINConv(nn)``: remove_self_loops; out = nn((1+eps)*x + sum_j x_j)."""
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    src, dst = src[keep], dst[keep]
    agg = torch.zeros_like(x)
    agg.index_add_(0, dst, x[src])
    eps = p[prefix + ".ginConv.eps"].to(x.dtype)
    h = (1 + eps) * x + agg
    return mlp(h, p, prefix + ".ginConv.nn", 3, "sigmoid", False, "train")


def forward(params, x, adj_e_index, adj_e_features, col_e_idx, *, depth,
            bn_mode="train", dtype=torch.float64, return_intermediates=False,
            edge_chunk=1 << 16, _capture_bn=False):
    """``TilinGNN.forward`` (TilinGNN.py:51-78).  Returns scores ``[N, 1]``.

    ``params``: reference state_dict (tensors, any float dtype); cast to ``dtype``.
    ``bn_mode``: "train" (the reference's behaviour, ml_solver.py:131) or "eval".
    """
    p = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in params.items()}
    x = x.to(dtype)
    adj_e_features = adj_e_features.to(dtype)
    adj_e_index = adj_e_index.long()
    col_e_idx = col_e_idx.long()
    inter = {}
    global _BN_CAPTURE
    _BN_CAPTURE = {} if _capture_bn else None

    h = mlp(x, p, "init_node_feature_trans", 2, "leaky", True, bn_mode)     # TilinGNN.py:54
    b1 = b2 = h
    middle = [h]
    for i in range(depth):
        # GraphConv: nnConv -> LeakyReLU -> BN   (edge_conv.py:24-30)
        pre1 = _leaky(nnconv_mean(b1, adj_e_index, adj_e_features, p,
                                  f"brch_1_graph_conv_layers.{i}", edge_chunk))
        g1 = batch_norm(pre1, p, f"brch_1_graph_conv_layers.{i}.batch_norm", bn_mode)
        # CollConv: ginConv -> LeakyReLU -> BN   (coll_conv.py:24-30)
        pre2 = _leaky(ginconv(b2, col_e_idx, p, f"brch_2_coll_conv_layers.{i}"))
        g2 = batch_norm(pre2, p, f"brch_2_coll_conv_layers.{i}.batch_norm", bn_mode)
        b2 = g2
        b1 = g1 * g2                                                       # TilinGNN.py:64
        if i - 2 >= 0:                                                     # residual_skip_num = 2
            b1 = b1 + middle[i - 2]
        middle.append(b1)
        if return_intermediates:
            inter[f"pre1_{i}"] = pre1
            inter[f"pre2_{i}"] = pre2
            inter[f"b1_{i}"] = b1
            inter[f"b2_{i}"] = b2
    z = torch.cat(middle, 1)                                               # TilinGNN.py:74
    z = mlp(z, p, "final_mlp.0", 4, "leaky", True, bn_mode)                # TilinGNN.py:45-46
    score = linear_trans(z, p, "final_mlp.1", "sigmoid", False, bn_mode)   # TilinGNN.py:47
    if _capture_bn:
        inter["bn_stats"], _BN_CAPTURE = _BN_CAPTURE, None
    if return_intermediates:
        inter["h0"] = h
        return score, inter
    return score


def calibrate_running_stats(params, x, adj_e_index, adj_e_features, col_e_idx, *, depth):
    """Copy of ``params`` whose BatchNorm running statistics are the (biased) batch statistics of a
    train-mode fp64 forward on this graph -- so that an eval-mode forward of a synthetic network is
    not saturated (with arbitrary running statistics it is: SURVEY.md §8c) and equals the train-mode
    one up to rounding."""
    _, it = forward(params, x, adj_e_index, adj_e_features, col_e_idx, depth=depth, bn_mode="train",
                    dtype=torch.float64, return_intermediates=True, _capture_bn=True)
    out = dict(params)
    for k, (mu, var) in it["bn_stats"].items():
        out[k + ".running_mean"] = mu.to(torch.float32)
        out[k + ".running_var"] = var.to(torch.float32)
    return out


def predict(params, node_feature, align_edge_index, align_edge_features,
            collide_edge_index, *, depth, bn_mode="train", dtype=torch.float32):
    """``ML_Solver.predict`` (ml_solver.py:29-49) with num_prob_maps = 1:
    all-ones when either edge set is empty, else column 0 of the forward."""
    import numpy as np
    n = node_feature.shape[0]
    if len(collide_edge_index) == 0 or len(align_edge_index) == 0:
        return np.ones(n, dtype=np.float32)
    s = forward(params,
                torch.as_tensor(node_feature).float(),
                torch.as_tensor(align_edge_index).long(),
                torch.as_tensor(align_edge_features).float(),
                torch.as_tensor(collide_edge_index).long(),
                depth=depth, bn_mode=bn_mode, dtype=dtype)
    return s[:, 0].to(torch.float32).numpy()


# --------------------------------------------------------------------------- #
# parameter helpers (shared by tests / bench so that oracle and CUDA path get  #
# IDENTICAL weights through the reference's state_dict keys)                   #
# --------------------------------------------------------------------------- #

def reference_param_shapes(d_x, d_e, depth, width=32):
    """Ordered {key: shape} of the reference state_dict (TilinGNN.py:14-48),
    including the aliased ``nnConv.nn.mlp.*`` entries and BN buffers."""
    F = width
    shapes = {}

    def lt(prefix, i, o, bn):
        shapes[prefix + ".linear.weight"] = (o, i)
        shapes[prefix + ".linear.bias"] = (o,)
        if bn:
            bnp = prefix + ".batch_norm"
            shapes[bnp + ".weight"] = (o,)
            shapes[bnp + ".bias"] = (o,)
            shapes[bnp + ".running_mean"] = (o,)
            shapes[bnp + ".running_var"] = (o,)
            shapes[bnp + ".num_batches_tracked"] = ()

    def bn(prefix, o):
        shapes[prefix + ".weight"] = (o,)
        shapes[prefix + ".bias"] = (o,)
        shapes[prefix + ".running_mean"] = (o,)
        shapes[prefix + ".running_var"] = (o,)
        shapes[prefix + ".num_batches_tracked"] = ()

    lt("init_node_feature_trans.mlp.0", d_x, F, True)
    lt("init_node_feature_trans.mlp.1", F, F, True)
    for i in range(depth):
        pre = f"brch_1_graph_conv_layers.{i}"
        dims = [d_e, 32, 64, F * F]
        for k in range(3):
            lt(f"{pre}.mlp.mlp.{k}", dims[k], dims[k + 1], False)
        shapes[pre + ".nnConv.root"] = (F, F)
        shapes[pre + ".nnConv.bias"] = (F,)
        for k in range(3):
            lt(f"{pre}.nnConv.nn.mlp.{k}", dims[k], dims[k + 1], False)
        bn(pre + ".batch_norm", F)
    for i in range(depth):
        pre = f"brch_2_coll_conv_layers.{i}"
        shapes[pre + ".ginConv.eps"] = (1,)
        dims = [F, 32, 64, F]
        for k in range(3):
            lt(f"{pre}.ginConv.nn.mlp.{k}", dims[k], dims[k + 1], False)
        bn(pre + ".batch_norm", F)
    dims = [F * (depth + 1), 256, 128, 64, F]
    for k in range(4):
        lt(f"final_mlp.0.mlp.{k}", dims[k], dims[k + 1], True)
    lt("final_mlp.1", F, 1, False)
    return shapes


def make_params(d_x, d_e, depth, width=32, seed=0, conditioned=True):
    """Seeded synthetic parameters under the reference's keys.

    ``conditioned=True`` draws weights so that every pre-BatchNorm tensor has a
    per-channel spread well above sqrt(BN_EPS) (the shipped checkpoints do not:
    SURVEY.md §7 "hard parts"), so the fp32-vs-fp64 1e-4 tolerance is
    meaningful on synthetic graphs.  ``conditioned=False`` is torch's default
    init of the reference modules (uniform +-1/sqrt(fan_in)).
    """
    g = torch.Generator().manual_seed(seed)
    shapes = reference_param_shapes(d_x, d_e, depth, width)
    p = {}
    for k, shp in shapes.items():
        if ".nnConv.nn.mlp." in k:
            continue                                    # alias, filled below
        if k.endswith("num_batches_tracked"):
            p[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            p[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("running_var"):
            p[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("batch_norm.weight"):
            p[k] = (0.75 + 0.5 * torch.rand(shp, generator=g)) if conditioned else torch.ones(shp)
        elif k.endswith("batch_norm.bias"):
            p[k] = (0.2 * torch.randn(shp, generator=g)) if conditioned else torch.zeros(shp)
        elif k.endswith("ginConv.eps"):
            p[k] = torch.zeros(1)
        elif k.endswith(".nnConv.root"):
            p[k] = (torch.rand(shp, generator=g) * 2 - 1) / (shp[0] ** 0.5)
        elif k.endswith(".nnConv.bias"):
            p[k] = (torch.rand(shp, generator=g) * 2 - 1) / (width ** 0.5)
        elif k.endswith("linear.weight"):
            fan_in = shp[1]
            gain = 1.0
            if conditioned and (".ginConv.nn.mlp." in k or ".mlp.mlp." in k):
                gain = 3.0                              # keep the sigmoid stacks out of their flat region
                if ".ginConv.nn.mlp.0." in k:
                    gain = 0.5                          # its input is a SUM over up to ~33 rows
            p[k] = gain * (torch.rand(shp, generator=g) * 2 - 1) / (fan_in ** 0.5)
        elif k.endswith("linear.bias"):
            p[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.1
        else:
            raise KeyError(k)
    for i in range(depth):                              # edge_conv.py:17-18 registers ONE MLP under two names
        pre = f"brch_1_graph_conv_layers.{i}"
        for k in range(3):
            for leaf in ("weight", "bias"):
                p[f"{pre}.nnConv.nn.mlp.{k}.linear.{leaf}"] = p[f"{pre}.mlp.mlp.{k}.linear.{leaf}"]
    return {k: p[k] for k in shapes}                    # reference key order
