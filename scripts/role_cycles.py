#!/usr/bin/env python
"""Where do the warp roles spend their cycles?  (TGNN_ROLE_DBG=1; CTA 0 of the last launch.)
  role_cycles.py [N] [DEG]   synthetic lattice: k_gin_w roles {cycles, wait for data, wait for buffer space} (and k_conv_t / z
                             when forced with TGNN_CONV)
  role_cycles.py bunny       604-node bunny layout, depth 20: the one-tile-per-CTA geometry of k_conv_h, per warp
                             {total, prologue, chunk loop, epilogue} cycles"""
import os, sys
os.environ["TGNN_ROLE_DBG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from tilingnn_b200 import TilinGNN, synthetic as syn
dev = torch.device("cuda:0")
small = len(sys.argv) > 1 and sys.argv[1] == "bunny"
if small:
    from _util import GOLDEN, load_ckpt, load_layout
    from tilingnn_b200.ml_solver import to_torch_tensor
    z = dict(np.load(os.path.join(GOLDEN, "c5_bunny.npz")))
    sg, graph = load_layout(z, "L0_")
    net = TilinGNN(int(z["d_e"]), 20, 32, node_features_dim=int(z["d_x"]))
    net.load_state_dict(load_ckpt("ckpt_30-60-90+equilateral.npz"), strict=True)
    net = net.to(dev).train()
    x, ai, af, ci, _ = to_torch_tensor(dev, sg.node_feature, sg.align_edge_index, sg.align_edge_features, sg.collide_edge_index)
    n = x.shape[0]
else:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    deg = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    torch.manual_seed(0)
    net = TilinGNN(19, 6, 32, node_features_dim=3).to(dev).train()
    x, ai, af, ci = syn.lattice_graph(n, deg, deg, 3, 19, seed=0, device=dev)
net.set_graph(n, ai, af, ci)
for _ in range(4):
    net.score(x)
torch.cuda.synchronize()
info = net.info()
print({k: info[k] for k in ("n_own", "conv_kernel", "gin_kernel", "t_rows", "t_blocks", "gin_window_tiles", "gin_direct_tiles", "adj_slots", "e_adj")})
names = ("total", "prologue", "chunk loop", "epilogue") if small else ("cycles", "wait0", "wait1", "wait2")
for k, rows in net.debug_role_cycles().items():
    print(k if not small else k.replace("k_conv_t|z", "k_conv_h (one tile per CTA / cluster)"))
    for w, r in enumerate(rows):
        if r[0]:
            print(f"  warp {w:2d}: " + "  ".join(f"{nm} {v:>10d}" + (f" ({100 * v / r[0]:5.1f}%)" if i else "") for i, (nm, v) in enumerate(zip(names, r))))
