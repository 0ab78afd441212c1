"""``tiling_a_region`` -- the driver of /root/reference/Tiling-Shape.py:28-86 without plotting:
load the complete graph and the checkpoint, crop candidate placements out of a silhouette, solve every
cropped layout with the greedy assembly, return (and optionally pickle) the solved layouts.

    python -m tilingnn_b200.tiling_shape --data data/30-60-90+equilateral --net pre-trained_models/30-60-90+equilateral.pth \\
           --silhouette silhouette/bunny.txt --out results/
"""
from __future__ import annotations

import argparse
import os
import pickle
import time

import numpy as np
import torch

from . import tile_graph_io as tio
from .ml_solver import ML_Solver
from .network import TilinGNN


def tiling_a_region(complete_graph_path, network_path, silhouette_path, device="cuda", network_depth=20,
                    network_width=32, tile_type_count=None, start_angle=0, end_angle=30, num_of_angle=1,
                    movement_delta_ratio=(0, 0.5), margin_padding_ratios=(0.5,), seed=2, out_dir=None, verbose=True,
                    solver_factory=None):
    """Defaults are the reference's (Tiling-Shape.py:52-54, inputs/config.py:38-45).  Returns a list of
    ``(solved_layout, score)``; ``solved_layout.predict`` is the 0/1 selection, ``.predict_order`` the order.
    ``solver_factory(complete_graph, state_dict)`` (tests) replaces the CUDA network + ``ML_Solver`` construction."""
    state = torch.load(network_path, map_location="cpu", weights_only=True)
    d_x = state["init_node_feature_trans.mlp.0.linear.weight"].shape[1]
    # environment.tile_count counts mirrored prototypes too (inputs/env.py:26-36), which the pickle alone cannot tell
    g = tio.load_complete_graph(complete_graph_path, d_x - 1 if tile_type_count is None else tile_type_count)
    if d_x != g.tile_type_count + 1:
        raise ValueError(f"checkpoint expects {d_x} node features, the complete graph gives {g.tile_type_count + 1}")
    if solver_factory is not None:
        solver = solver_factory(g, state)
    else:
        device = torch.device(device)
        network = TilinGNN(adj_edge_features_dim=g.total_feature_dim, network_depth=network_depth,
                           network_width=network_width, node_features_dim=d_x).to(device)
        solver = ML_Solver(None, device, g, network, num_prob_maps=1)
        solver.load_saved_network(network_path)
    exterior, interiors = tio.load_polygons(silhouette_path)
    layouts = tio.crop_multiple_layouts_from_contour(exterior, interiors, g, start_angle=start_angle, end_angle=end_angle,
                                                     num_of_angle=num_of_angle, movement_delta_ratio=movement_delta_ratio,
                                                     margin_padding_ratios=margin_padding_ratios)
    rng = np.random.RandomState(seed)
    solutions = []
    for idx, layout in enumerate(layouts):
        t0 = time.perf_counter()
        solved, score = solver.solve(layout, rng=rng)
        dt = time.perf_counter() - t0
        solutions.append((solved, score))
        if verbose:
            print(f"layout {idx}: {layout.node_feature.shape[0]} candidates, {int(solved.predict.sum())} tiles placed in "
                  f"{solved.greedy_rounds} rounds, score {score:.6f}, {dt:.3f} s")
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)
            with open(os.path.join(out_dir, f"{score}_{idx}_data.pkl"), "wb") as f:      # write_bricklayout(with_features=False)
                pickle.dump({"tiles": solved.tiles, "predict": solved.predict, "predict_order": solved.predict_order,
                             "predict_probs": solved.predict_probs, "score": score}, f)
    return solutions


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True, help="directory holding complete_graph_ring<k>.pkl")
    ap.add_argument("--ring", type=int, default=9)
    ap.add_argument("--net", required=True)
    ap.add_argument("--silhouette", required=True)
    ap.add_argument("--out", default=None)
    ap.add_argument("--device", default="cuda")
    a = ap.parse_args()
    tiling_a_region(os.path.join(a.data, f"complete_graph_ring{a.ring}.pkl"), a.net, a.silhouette, a.device, out_dir=a.out)


if __name__ == "__main__":
    main()
