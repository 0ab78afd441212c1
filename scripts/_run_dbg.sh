python - <<'PY'
import os, sys
os.environ["TGNN_ROLE_DBG"] = "1"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from _util import GOLDEN, load_ckpt, load_layout
from tilingnn_b200 import TilinGNN
from tilingnn_b200.ml_solver import to_torch_tensor
dev = torch.device("cuda:0")
z = dict(np.load(os.path.join(GOLDEN, "c5_bunny.npz")))
sg, graph = load_layout(z, "L0_")
net = TilinGNN(int(z["d_e"]), 20, 32, node_features_dim=int(z["d_x"]))
net.load_state_dict(load_ckpt("ckpt_30-60-90+equilateral.npz"), strict=True)
net = net.to(dev).train()
x, ai, af, ci, _ = to_torch_tensor(dev, sg.node_feature, sg.align_edge_index, sg.align_edge_features, sg.collide_edge_index)
net.set_graph(x.shape[0], ai, af, ci)
for _ in range(4): net.score(x)
torch.cuda.synchronize()
print(net.info())
for k, rows in net.debug_role_cycles().items():
    print(k)
    for w, r in enumerate(rows):
        if r[0]: print(f"  warp {w:2d}: total {r[0]:>8d}  prologue {r[1]:>8d}  loop {r[2]:>8d}  epilogue {r[3]:>8d}")
PY
