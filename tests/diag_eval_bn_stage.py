#!/usr/bin/env python
"""Which stage carries the eval-BN error of config 1 (VERDICT r1 'weak' #2)?  Runs the shipped 30-60-90 checkpoint on the
heart crop / complete graph in eval mode, stops after every layer and compares pre1, pre2 (before BatchNorm), g2 and b1
with the fp64 oracle's intermediates -- next to the gain gamma/sqrt(running_var + eps) each BatchNorm applies.
(Diagnostic; uses the oracle as the checker.  Run on the GPU box.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import load_ckpt, load_graph                      # noqa: E402
from oracle import tilingnn_oracle as orc                     # noqa: E402
from tilingnn_b200 import TilinGNN                            # noqa: E402

dev = torch.device("cuda:0")
for graph in (sys.argv[1:] or ["c1_heart.npz"]):
    z, x, ai, af, ci = load_graph(graph)
    ckpt = load_ckpt()
    score, inter = orc.forward(ckpt, x, ai, af, ci, depth=20, bn_mode="eval", dtype=torch.float64, return_intermediates=True)
    s32 = orc.forward(ckpt, x, ai, af, ci, depth=20, bn_mode="eval", dtype=torch.float32)
    net = TilinGNN(19, 20, 32, node_features_dim=3)
    net.load_state_dict(ckpt, strict=True)
    net = net.to(dev).eval()
    args = dict(x=x.to(dev), adj_e_index=ai.to(dev), adj_e_features=af.to(dev), col_e_idx=ci.to(dev))
    print(f"== {graph}: N={x.shape[0]}  (abs errors vs fp64 oracle; gain = max_c |gamma_c| / sqrt(running_var_c + 1e-5))")
    print("layer  pre1      pre2      g2        b1        gainA     gainC")
    for i in range(20):
        net.debug_set_stop_layer(i)
        net(**args)
        err = {}
        for name, key in (("pre1", f"pre1_{i}"), ("pre2", f"pre2_{i}"), ("g2", f"b2_{i}"), (f"mid_{i + 1}", f"b1_{i}")):
            got = net.debug_read(name).double().cpu()
            err[name] = float((got - inter[key]).abs().max())
        gain = []
        for br in ("brch_1_graph_conv_layers", "brch_2_coll_conv_layers"):
            w, v = ckpt[f"{br}.{i}.batch_norm.weight"].double(), ckpt[f"{br}.{i}.batch_norm.running_var"].double()
            gain.append(float((w.abs() / torch.sqrt(v + 1e-5)).max()))
        print(f"{i:4d}   {err['pre1']:.2e}  {err['pre2']:.2e}  {err['g2']:.2e}  {err[f'mid_{i + 1}']:.2e}  {gain[0]:8.1f}  {gain[1]:8.1f}")
    net.debug_set_stop_layer(-1)
    s = net(**args)[0][:, 0].double().cpu()
    print(f"scores: ours {float((s - score[:, 0]).abs().max()):.2e}   oracle fp32 {float((s32[:, 0].double() - score[:, 0]).abs().max()):.2e}")
