"""Per-call latency breakdown of ML_Solver.predict on a real-size layout (config 5, ~600 nodes, depth 20)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import GOLDEN, load_ckpt, load_layout
from tilingnn_b200 import ML_Solver, TilinGNN
from tilingnn_b200.ml_solver import to_torch_tensor
z = dict(np.load(os.path.join(GOLDEN, "c5_bunny.npz")))
sg, graph = load_layout(z, "L0_")
dev = torch.device("cuda:0")
net = TilinGNN(int(z["d_e"]), 20, 32, node_features_dim=int(z["d_x"]))
net.load_state_dict(load_ckpt("ckpt_30-60-90+equilateral.npz"), strict=True)
net = net.to(dev).train()
solver = ML_Solver(None, dev, graph, net, 1)
for _ in range(3): solver.predict(sg)
def T(f, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
ms, _ = T(lambda: solver.predict(sg)); print(f"predict            {ms:8.3f} ms")
ms, t = T(lambda: to_torch_tensor(dev, sg.node_feature, sg.align_edge_index, sg.align_edge_features, sg.collide_edge_index)); print(f"to_torch_tensor    {ms:8.3f} ms")
x, ai, af, ci, _ = t
ms, _ = T(lambda: net.set_graph(x.shape[0], ai, af, ci)); print(f"set_graph          {ms:8.3f} ms")
for _ in range(3): net.score(x)          # (the third call captures the CUDA graph)
ms, _ = T(lambda: net.score(x)); print(f"score (forward)    {ms:8.3f} ms   launches {net.info()['launches_per_forward']}")
net.set_profiling(True); net.score(x); print({k: round(v[0], 3) for k, v in net.profile().items()}); net.set_profiling(False)
ms, _ = T(lambda: net.score(x).cpu()); print(f"score + D2H        {ms:8.3f} ms")
