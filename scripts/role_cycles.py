#!/usr/bin/env python
"""Where do the warp roles of k_conv_t / k_gin_w spend their cycles?  (TGNN_ROLE_DBG=1; CTA 0 of the last launch.)"""
import os, sys
os.environ["TGNN_ROLE_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tilingnn_b200 import TilinGNN, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
deg = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = TilinGNN(19, 6, 32, node_features_dim=3).to(dev).train()
x, ai, af, ci = syn.lattice_graph(n, deg, deg, 3, 19, seed=0, device=dev)
net.set_graph(n, ai, af, ci)
for _ in range(3):
    net.score(x)
torch.cuda.synchronize()
info = net.info()
print({k: info[k] for k in ("conv_kernel", "gin_kernel", "t_rows", "t_blocks", "gin_window_tiles", "gin_direct_tiles", "adj_slots", "e_adj")})
for k, rows in net.debug_role_cycles().items():
    print(k)
    for w, r in enumerate(rows):
        if r[0]:
            print(f"  warp {w:2d}: cycles {r[0]:>10d}  wait0 {r[1]:>10d} ({100 * r[1] / r[0]:5.1f}%)  wait1 {r[2]:>10d} ({100 * r[2] / r[0]:5.1f}%)  wait2 {r[3]:>10d} ({100 * r[3] / r[0]:5.1f}%)")
