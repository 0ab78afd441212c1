"""Forward latency on a real-size layout (config 5 bunny crop, ~600 nodes, depth 20) or a small synthetic lattice.
usage: small_forward.py [--reps R] [--lattice N DEG] [--eager]     (run under ncu for the per-kernel launch list)"""
import argparse, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--lattice", type=int, nargs=2, default=None)
ap.add_argument("--eager", action="store_true")
args = ap.parse_args()
if args.eager: os.environ["TGNN_GRAPH"] = "0"
from _util import GOLDEN, load_ckpt, load_layout
from tilingnn_b200 import TilinGNN
from tilingnn_b200.ml_solver import to_torch_tensor
dev = torch.device("cuda:0")
if args.lattice:
    from tilingnn_b200 import synthetic as syn
    n, deg = args.lattice
    x, ai, af, ci = [t.to(dev) for t in syn.lattice_graph(n, deg, deg, seed=0)]
    torch.manual_seed(0)
    net = TilinGNN(19, 6, 32, node_features_dim=3)
    net = net.to(dev).train()
else:
    z = dict(np.load(os.path.join(GOLDEN, "c5_bunny.npz")))
    sg, graph = load_layout(z, "L0_")
    net = TilinGNN(int(z["d_e"]), 20, 32, node_features_dim=int(z["d_x"]))
    net.load_state_dict(load_ckpt("ckpt_30-60-90+equilateral.npz"), strict=True)
    net = net.to(dev).train()
    x, ai, af, ci, _ = to_torch_tensor(dev, sg.node_feature, sg.align_edge_index, sg.align_edge_features, sg.collide_edge_index)
net.set_graph(x.shape[0], ai, af, ci)
for _ in range(4): s = net.score(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(args.reps): s = net.score(x)
b.record(); torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / args.reps * 1e3
print(f"nodes {x.shape[0]}  forward: {a.elapsed_time(b) / args.reps:.4f} ms (device), {wall:.4f} ms (wall)  launches {net.info()['launches_per_forward']}"
      f"  conv_kernel {net.info()['conv_kernel']} gin_kernel {net.info().get('gin_kernel')}  checksum {float(s.double().sum()):.9f}")
