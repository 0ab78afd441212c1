#!/usr/bin/env python
"""BASELINE.json config 5: 30-60-90+equilateral, bunny.txt, the four layouts of Tiling-Shape.py:52-54 --
scoring + greedy assembly wall-clock, this repo's CUDA path next to the CPU port of the reference (the oracle
network, fp32, all host threads, driving the same greedy loop).  Inputs: tests/golden/c5_bunny.npz and the
shipped checkpoint (tests/golden/ckpt_30-60-90+equilateral.npz).  Prints ONE JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _util import GOLDEN, load_ckpt, load_layout            # noqa: E402
from oracle import tilingnn_oracle as orc                   # noqa: E402   (cpu_baseline leg only)
from tilingnn_b200 import ML_Solver, TilinGNN, greedy       # noqa: E402


class OracleSolver:
    def __init__(self, params, graph):
        self.params, self.complete_graph, self.calls, self.nodes = params, graph, 0, 0

    def predict(self, lay):
        n = lay.node_feature.shape[0]
        if np.size(lay.collide_edge_index) == 0 or np.size(lay.align_edge_index) == 0:
            return np.ones(n, dtype=np.float32)
        self.calls += 1
        self.nodes += n
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        s = orc.forward(self.params, t(lay.node_feature, torch.float32), t(lay.align_edge_index, torch.long),
                        t(lay.align_edge_features, torch.float32), t(lay.collide_edge_index, torch.long), depth=20,
                        bn_mode="train", dtype=torch.float32)
        return s[:, 0].float().numpy()


def main():
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    z = dict(np.load(os.path.join(GOLDEN, "c5_bunny.npz")))
    ckpt = load_ckpt("ckpt_30-60-90+equilateral.npz")
    layouts = [load_layout(z, prefix=f"L{i}_") for i in range(int(z["n_layouts"]))]
    dev = torch.device("cuda:0")
    net = TilinGNN(int(z["d_e"]), 20, 32, node_features_dim=int(z["d_x"]))
    net.load_state_dict(ckpt, strict=True)
    net = net.to(dev).train()
    calls = {"n": 0, "nodes": 0}
    solver = ML_Solver(None, dev, None, net, 1)             # one solver for all layouts, as Tiling-Shape.py:37

    def run_gpu(seed):
        out = []
        rng = np.random.RandomState(seed)
        for sg, graph in layouts:
            solver.complete_graph = graph
            solved, score = solver.solve(sg, rng=rng)
            calls["n"] += solved.greedy_rounds + 1
            out.append((int(solved.predict.sum()), score))
        return out
    run_gpu(0)                                              # warm-up: library load, parameter upload
    torch.cuda.synchronize()
    times = []
    for r in range(repeats):
        calls["n"] = 0
        t0 = time.perf_counter()
        res = run_gpu(2)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    gpu_s = float(np.median(times))

    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    rng = np.random.RandomState(2)
    cpu_res, cpu_calls = [], 0
    for sg, graph in layouts:
        s = OracleSolver(ckpt, graph)
        r = greedy.solve_by_probablistic_greedy(s, sg, rng=rng)
        s.predict(sg)                                       # ML_Solver.solve's final scoring pass (ml_solver.py:65)
        cpu_calls += s.calls
        cpu_res.append((int(r.selection.sum()), r.score))
    cpu_s = time.perf_counter() - t0
    print(json.dumps({
        "metric": "Tiling-Shape scoring + greedy assembly wall-clock (config 5)", "unit": "s", "higher_is_better": False,
        "value": gpu_s, "times": times, "network_calls": calls["n"],
        "layouts": [{"nodes": int(sg.node_feature.shape[0]), "tiles_placed": a, "score": b} for (sg, _), (a, b) in zip(layouts, res)],
        "cpu_baseline": {"value": cpu_s, "unit": "s", "cores": os.cpu_count(), "kind": "port", "network_calls": cpu_calls,
                         "sample": "the same 4 layouts, oracle network fp32 driving the same greedy loop (one run)",
                         "layouts": [{"tiles_placed": a, "score": b} for a, b in cpu_res]},
        "speedup_vs_cpu_port": cpu_s / gpu_s,
        "config": {"workload": "30-60-90+equilateral, bunny.txt, 4 layouts (604/562/591/565 candidate tiles), depth 20, "
                               "train-mode BatchNorm, shipped checkpoint"}}))


if __name__ == "__main__":
    main()
