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
