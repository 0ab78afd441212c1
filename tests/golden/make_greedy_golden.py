"""Golden vectors for the greedy assembly / sub-layout / loss callers (run in the build container only).

    python tests/golden/make_greedy_golden.py

Needs /root/reference.  The expected outputs are produced by the reference's OWN source text: the functions
``solve_by_probablistic_greedy``, ``label_collision_neighbor`` and class ``SelectionSolution`` of
util/algorithms.py, the method ``BrickLayout.compute_sub_layout`` of tiling/brick_layout.py and
``Losses.calculate_unsupervised_loss`` of solver/ml_solver/losses.py are cut out of the files with ``ast`` and
executed unmodified.  Their modules cannot be imported whole here (shapely, matplotlib, PyQt5 are absent), so the
only stand-ins are: ``Polygon`` (the running union polygon of SelectionSolution, irrelevant to the greedy's
decisions), ``create_solution`` (returns the labelled nodes; the score needs shapely) and ``inputs.config``'s three
loss weights.  The network is replaced by the deterministic ``fake_predict`` of tests/_util.py so that the golden
run and the test make bit-identical decisions.

Writes greedy_heart.npz: the heart crop's arrays, the selection / order / rounds for numpy seed 2, the sub-layout
of the first round boundary, and the unsupervised losses of two probability maps.
"""
import ast
import math
import os
import sys
import types
from collections import OrderedDict, defaultdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tilingnn_b200 import tile_graph_io as tio            # noqa: E402
from _util import fake_predict                            # noqa: E402

REF = "/root/reference"


def cut(path, names):
    """Source text of the top-level defs / classes (or Class.method) called ``names``."""
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            out[node.name] = ast.get_source_segment(src, node)
        if isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and f"{node.name}.{sub.name}" in names:
                    import textwrap
                    seg = ast.get_source_segment(src, sub)
                    out[f"{node.name}.{sub.name}"] = textwrap.dedent(" " * sub.col_offset + seg)
    return out


class Polygon:                       # stand-in: SelectionSolution only unions tile polygons into it
    def __init__(self, *a, **k):
        pass

    def union(self, other):
        return self

    def buffer(self, *_):
        return self


class _Tile:
    tile_poly = Polygon()


class _Graph:
    tiles = defaultdict(_Tile)


def pack_layout(g, sg, prefix=""):
    """Arrays of one cropped layout + what solution_score needs of the complete graph (rings padded with NaN)."""
    uniq, inv_f = np.unique(np.asarray(sg.align_edge_features, dtype=np.float64), axis=0, return_inverse=True)
    rings = [np.asarray(g.tile_rings[t]) for t in sg.tiles]
    m = max(len(r) for r in rings)
    pad = np.full((len(rings), m, 2), np.nan)
    for i, r in enumerate(rings):
        pad[i, :len(r)] = r
    d = dict(node_feature=sg.node_feature, tiles=sg.tiles, collide_edge_index=sg.collide_edge_index.astype(np.int32),
             align_edge_index=sg.align_edge_index.astype(np.int32), align_feat_rows=uniq,
             align_feat_id=inv_f.reshape(-1).astype(np.int16), tile_rings=pad.astype(np.float64),
             max_area=g.max_area, max_align_length=g.max_align_length)
    return {prefix + k: v for k, v in d.items()}


def config5():
    """config 5 of BASELINE.json: 30-60-90+equilateral, bunny.txt, the four layouts of Tiling-Shape.py:52-54, and the
    shipped checkpoint (fp32, aliased keys dropped)."""
    env = "30-60-90+equilateral"
    g = tio.load_complete_graph(os.path.join(REF, f"data/{env}/complete_graph_ring9.pkl"), tile_type_count=4)   # 2 tiles x mirror (env.py:26-36)
    ext, ints = tio.load_polygons(os.path.join(REF, "silhouette/bunny.txt"))
    layouts = tio.crop_multiple_layouts_from_contour(ext, ints, g, start_angle=0, end_angle=30, num_of_angle=1,
                                                     movement_delta_ratio=[0, 0.5], margin_padding_ratios=[0.5])
    out = {"n_layouts": len(layouts), "d_e": g.total_feature_dim, "d_x": g.tile_type_count + 1}
    for i, sg in enumerate(layouts):
        out.update(pack_layout(g, sg, prefix=f"L{i}_"))
    sd = torch.load(os.path.join(REF, f"pre-trained_models/{env}.pth"), map_location="cpu", weights_only=True)
    # scores of the reference's own graph_networks code (oracle/ref_harness.py) on layout 0: a second tile set
    # (D_x = 5, D_e = 66, 41 edge types) for the network parity tests
    from oracle import ref_harness as rh
    sg = layouts[0]
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
    for mode in ("train", "eval"):
        for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
            sc = rh.run_reference(sd, t(sg.node_feature, dt), t(sg.align_edge_index, torch.long), t(sg.align_edge_features, dt),
                                  t(sg.collide_edge_index, torch.long), depth=20, bn_mode=mode, dtype=dt)
            out[f"L0_ref_{mode}_{name}"] = sc[:, 0].double().numpy()
    np.savez_compressed(os.path.join(HERE, "c5_bunny.npz"), **out)
    np.savez_compressed(os.path.join(HERE, f"ckpt_{env}.npz"),
                        **{k: v.float().numpy() for k, v in sd.items()
                           if ".nnConv.nn.mlp." not in k and not k.endswith("num_batches_tracked")})
    print(f"c5_bunny.npz: {[l.node_feature.shape[0] for l in layouts]} nodes, d_x={out['d_x']} d_e={out['d_e']}")


def third_tile_set():
    """45-45-90+rectangle (symmetry_tiles = False: D_x = 3, D_e = 22), heart.txt layout 0, shipped checkpoint: scores of the
    reference's own graph_networks code -- with the 30-60-90 and 30-60-90+equilateral files this covers all three shipped
    checkpoints."""
    from oracle import ref_harness as rh
    env = "45-45-90+rectangle"
    sd = torch.load(os.path.join(REF, f"pre-trained_models/{env}.pth"), map_location="cpu", weights_only=True)
    d_x = sd["init_node_feature_trans.mlp.0.linear.weight"].shape[1]
    g = tio.load_complete_graph(os.path.join(REF, f"data/{env}/complete_graph_ring9.pkl"), tile_type_count=d_x - 1)
    ext, ints = tio.load_polygons(os.path.join(REF, "silhouette/heart.txt"))
    sg = tio.crop_multiple_layouts_from_contour(ext, ints, g, start_angle=0, end_angle=30, num_of_angle=1,
                                                movement_delta_ratio=[0, 0.5], margin_padding_ratios=[0.5])[0]
    out = {"d_x": d_x, "d_e": g.total_feature_dim}
    out.update(pack_layout(g, sg))
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
    for mode in ("train", "eval"):
        for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
            sc = rh.run_reference(sd, t(sg.node_feature, dt), t(sg.align_edge_index, torch.long), t(sg.align_edge_features, dt),
                                  t(sg.collide_edge_index, torch.long), depth=20, bn_mode=mode, dtype=dt)
            out[f"ref_{mode}_{name}"] = sc[:, 0].double().numpy()
    np.savez_compressed(os.path.join(HERE, "c1_rect_heart.npz"), **out)
    np.savez_compressed(os.path.join(HERE, f"ckpt_{env}.npz"),
                        **{k: v.float().numpy() for k, v in sd.items()
                           if ".nnConv.nn.mlp." not in k and not k.endswith("num_batches_tracked")})
    print(f"c1_rect_heart.npz: N={sg.node_feature.shape[0]} d_x={d_x} d_e={g.total_feature_dim} "
          f"fp32-vs-fp64 train {np.abs(out['ref_train_f32'] - out['ref_train_f64']).max():.2e} "
          f"eval {np.abs(out['ref_eval_f32'] - out['ref_eval_f64']).max():.2e}")


def main():
    third_tile_set()
    config5()
    g = tio.load_complete_graph(os.path.join(REF, "data/30-60-90/complete_graph_ring9.pkl"))
    ext, ints = tio.load_polygons(os.path.join(REF, "silhouette/heart.txt"))
    sg = tio.crop_multiple_layouts_from_contour(ext, ints, g, start_angle=0, end_angle=30, num_of_angle=1,
                                                movement_delta_ratio=[0, 0.5], margin_padding_ratios=[0.5])[0]
    # ---- the reference's code, cut out of its files ---------------------------------------------
    alg = cut(os.path.join(REF, "util/algorithms.py"),
              {"solve_by_probablistic_greedy", "label_collision_neighbor", "SelectionSolution"})
    bl = cut(os.path.join(REF, "tiling/brick_layout.py"), {"BrickLayout.compute_sub_layout"})
    ls = cut(os.path.join(REF, "solver/ml_solver/losses.py"), {"Losses.calculate_unsupervised_loss"})

    class BrickLayout:
        def __init__(self, complete_graph, node_feature, collide_edge_index, collide_edge_features, align_edge_index,
                     align_edge_features, re_index, target_polygon=None):
            self.complete_graph, self.node_feature = complete_graph, node_feature
            self.collide_edge_index, self.collide_edge_features = collide_edge_index, collide_edge_features
            self.align_edge_index, self.align_edge_features = align_edge_index, align_edge_features
            self.re_index, self.target_polygon = re_index, target_polygon
            self.inverse_index = defaultdict(int)
            for k, v in re_index.items():
                self.inverse_index[v] = k
    ns_bl = {"np": np, "BrickLayout": BrickLayout}
    exec(bl["BrickLayout.compute_sub_layout"], ns_bl)
    BrickLayout.compute_sub_layout = ns_bl["compute_sub_layout"]

    captured = {}

    def create_solution(new_predict, origin_layout):
        captured["labelled"] = OrderedDict(new_predict.labelled_nodes)
        sel = np.zeros(origin_layout.node_feature.shape[0])
        order = [k for k, v in new_predict.labelled_nodes.items() if v == 1]
        sel[order] = 1
        return float("nan"), sel, order
    ns = {"np": np, "OrderedDict": OrderedDict, "Polygon": Polygon, "EPS": 1e-7, "create_solution": create_solution}
    for k in ("SelectionSolution", "label_collision_neighbor", "solve_by_probablistic_greedy"):
        exec(alg[k], ns)

    rounds = []

    class FakeSolver:
        def predict(self, layout):
            ci = np.asarray(layout.collide_edge_index)
            ai = np.asarray(layout.align_edge_index)
            rounds.append(layout.node_feature.shape[0])
            if len(ci) == 0 or len(ai) == 0:                      # ml_solver.py:31-32
                return np.ones(layout.node_feature.shape[0], dtype=np.float32)
            return fake_predict(layout.node_feature, ci, ai)

    layout = BrickLayout(_Graph(), sg.node_feature, sg.collide_edge_index, sg.collide_edge_features,
                         sg.align_edge_index, sg.align_edge_features, {int(t): i for i, t in enumerate(sg.tiles)})
    np.random.seed(2)
    selection, _, order = ns["solve_by_probablistic_greedy"](FakeSolver(), layout)

    # one sub-layout, straight from the reference's method: drop every third node
    pred = ns["SelectionSolution"](sg.node_feature.shape[0])
    for i in range(0, sg.node_feature.shape[0], 3):
        pred.labelled_nodes[i] = 0
        pred.unlabelled_nodes.pop(i)
    sub, inv = layout.compute_sub_layout(pred)

    # unsupervised loss of two probability maps
    cfg = types.SimpleNamespace(COLLISION_WEIGHT=1 / math.log(1 + 1e-1), ALIGN_LENGTH_WEIGHT=0.02, AVG_AREA_WEIGHT=1)
    ns_l = {"torch": torch, "np": np, "math": math, "time": __import__("time"), "config": cfg, "eps": 1e-7}
    exec(ls["Losses.calculate_unsupervised_loss"].replace("@staticmethod\n", ""), ns_l)
    rng = np.random.RandomState(0)
    probs = rng.uniform(0.01, 0.99, size=(sg.node_feature.shape[0], 2))
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
    loss, min_index, losses = ns_l["calculate_unsupervised_loss"](
        t(probs, torch.float64), t(sg.node_feature, torch.float64), t(sg.collide_edge_index, torch.long),
        t(sg.align_edge_index, torch.long), t(sg.align_edge_features, torch.float64))

    np.savez_compressed(
        os.path.join(HERE, "greedy_heart.npz"), **pack_layout(g, sg),
        selection=selection, order=np.asarray(order), round_sizes=np.asarray(rounds),
        labelled_keys=np.asarray(list(captured["labelled"].keys())), labelled_vals=np.asarray(list(captured["labelled"].values())),
        sub_keep=np.asarray([inv[i] for i in range(len(inv))]), sub_node_feature=sub.node_feature,
        sub_collide=np.asarray(sub.collide_edge_index).astype(np.int32), sub_align=np.asarray(sub.align_edge_index).astype(np.int32),
        sub_align_feat_col1=np.asarray(sub.align_edge_features)[:, 1],
        loss_probs=probs, losses=np.asarray(losses), loss_min_index=int(min_index))
    print(f"greedy_heart.npz: N={sg.node_feature.shape[0]} selected={int(selection.sum())} rounds={len(rounds)} "
          f"round sizes {rounds[:8]}... losses {losses}")


if __name__ == "__main__":
    main()
