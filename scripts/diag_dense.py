"""A/B of the tcgen05 dense stage against the CUDA-core one (TGNN_DENSE=ffma) on the same weights."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import tilingnn_oracle as orc
from tilingnn_b200 import TilinGNN, synthetic as syn
os.environ["TGNN_CHECK"] = "1"
dev = torch.device("cuda:0")
for n, deg, depth in ((300, 8, 2), (5000, 8, 6), (20000, 32, 6)):
    x, ai, af, ci = syn.lattice_graph(n, deg, deg, seed=0)
    p = orc.make_params(3, 19, depth, seed=0)
    gold = orc.forward(p, x, ai, af, ci, depth=depth, dtype=torch.float64)[:, 0]
    res = {}
    for mode in ("ffma", "tc"):
        if mode == "ffma": os.environ["TGNN_DENSE"] = "ffma"
        else: os.environ.pop("TGNN_DENSE", None)
        net = TilinGNN(19, depth, 32, node_features_dim=3); net.load_state_dict(p); net = net.to(dev).train()
        try:
            s = net(x=x.to(dev), adj_e_index=ai.to(dev), adj_e_features=af.to(dev), col_e_idx=ci.to(dev))[0][:, 0].double().cpu()
            res[mode] = s
            print(f"N={n} deg={deg} L={depth} {mode}: err vs fp64 {float((s - gold).abs().max()):.3e} finite={bool(torch.isfinite(s).all())}")
        except Exception as e:
            print(f"N={n} {mode}: EXCEPTION {e}")
    if len(res) == 2:
        print("   tc vs ffma:", float((res['tc'] - res['ffma']).abs().max()))
