"""Helpers shared by the tests: golden-vector loading (tests/golden/*.npz, made by make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def with_aliases(params, depth):
    """Re-create the aliased nnConv.nn.mlp.* keys (edge_conv.py:17-18 registers one MLP twice)."""
    out = dict(params)
    for i in range(depth):
        pre = f"brch_1_graph_conv_layers.{i}"
        for k in range(3):
            for leaf in ("weight", "bias"):
                out[f"{pre}.nnConv.nn.mlp.{k}.linear.{leaf}"] = out[f"{pre}.mlp.mlp.{k}.linear.{leaf}"]
    return out


def load_ckpt(name="ckpt_30-60-90.npz", depth=20):
    z = np.load(os.path.join(GOLDEN, name))
    return with_aliases({k: torch.from_numpy(z[k]) for k in z.files}, depth)


def load_graph(name):
    """Returns (dict of golden arrays, x, adj_index, adj_feat, col_index) as torch tensors."""
    z = dict(np.load(os.path.join(GOLDEN, name)))
    x = torch.from_numpy(z["x"]).float()
    ai = torch.from_numpy(z["adj_index"]).long()
    ci = torch.from_numpy(z["col_index"]).long()
    if "adj_feat" in z:
        af = torch.from_numpy(z["adj_feat"]).float()
    else:
        af = torch.from_numpy(z["adj_feat_rows"][z["adj_feat_id"].astype(np.int64)]).float()
    return z, x, ai, af, ci


def syn_small_params(z):
    depth = int(z["depth"])
    p = {k[len("param:"):]: torch.from_numpy(v) for k, v in z.items() if k.startswith("param:")}
    return with_aliases(p, depth), depth


def fake_predict(node_feature, collide_edge_index, align_edge_index):
    """Deterministic stand-in for the network in the greedy-assembly tests: a probability per node computed from
    the arrays the reference and this repo both hand to ``predict`` (shared with make_greedy_golden.py)."""
    n = node_feature.shape[0]
    ci = np.asarray(collide_edge_index).reshape(2, -1).astype(np.int64)
    ai = np.asarray(align_edge_index).reshape(2, -1).astype(np.int64)
    dc = np.bincount(ci[1], minlength=n).astype(np.float64)
    da = np.bincount(ai[1], minlength=n).astype(np.float64)
    i = np.arange(n, dtype=np.float64)
    v = np.sin(dc * 0.37 + da * 1.13 + i * 0.071 + n * 0.013 + np.asarray(node_feature, dtype=np.float64)[:, 0]) * 1e3
    return (0.05 + 0.9 * (v - np.floor(v))).astype(np.float32)


def load_layout(z, prefix=""):
    """(SuperGraph, complete-graph stand-in) from arrays written by make_greedy_golden.pack_layout.  The stand-in
    carries what Losses.solution_score reads: tile rings by complete-graph index, max_area, max_align_length."""
    from types import SimpleNamespace
    from tilingnn_b200.tile_graph_io import SuperGraph
    g = lambda k: z[prefix + k]
    af = g("align_feat_rows")[g("align_feat_id").astype(np.int64)]
    ci = g("collide_edge_index").astype(np.int64)
    sg = SuperGraph(g("node_feature"), ci, np.zeros((ci.shape[1], af.shape[1])), g("align_edge_index").astype(np.int64),
                    af, g("tiles"))
    rings = {int(t): r[~np.isnan(r[:, 0])] for t, r in zip(g("tiles"), g("tile_rings"))}
    graph = SimpleNamespace(tile_rings=rings, max_area=float(g("max_area")), max_align_length=float(g("max_align_length")))
    return sg, graph
