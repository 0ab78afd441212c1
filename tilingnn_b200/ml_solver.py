"""``ML_Solver`` -- the caller side of the scoring path, mirroring
/root/reference/solver/ml_solver/ml_solver.py:13-81,129-136 (``predict``, ``get_predict_probs``,
``load_saved_network``) and ``get_network_prediction``
(/root/reference/graph_networks/network_utils.py:4-21) for the ``tilingnn_b200.TilinGNN`` module.

A "layout" is anything with the five numpy attributes of the reference's ``BrickLayout``
(/root/reference/tiling/brick_layout.py:22-31): ``node_feature``, ``align_edge_index``,
``align_edge_features``, ``collide_edge_index``, ``collide_edge_features`` --
``tilingnn_b200.tile_graph_io.SuperGraph`` is one.
"""
from __future__ import annotations

import traceback
from copy import copy, deepcopy

import numpy as np
import torch


def get_network_prediction(network, x, adj_e_index, adj_e_features, col_e_idx, col_e_features=None):
    """network_utils.py:4-21: keyword call, print the traceback and re-raise, return probs only."""
    try:
        probs, *_ = network(x=x, adj_e_index=adj_e_index, adj_e_features=adj_e_features,
                            col_e_idx=col_e_idx, col_e_features=col_e_features)
    except Exception:
        print(traceback.format_exc())
        raise
    return probs


def to_torch_tensor(device, node_feature, align_edge_index, align_edge_features, collide_edge_index,
                    collide_edge_features=None):
    """util/data_util.py:110-117 -- except that the collision edge FEATURES, which the network never
    reads (TilinGNN.py:51,63), are not uploaded (9.3 MB of the reference's 16 MB per call on the
    30-60-90 complete graph)."""
    pin = device.type == "cuda"

    def up(a, dt):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        if pin and t.numel() > 0:
            t = t.pin_memory()
        return t.to(device, non_blocking=True)
    return (up(node_feature, torch.float32), up(align_edge_index, torch.int64),
            up(align_edge_features, torch.float32), up(collide_edge_index, torch.int64), None)


class ML_Solver:
    def __init__(self, debugger, device, complete_graph, network, num_prob_maps=1):
        if num_prob_maps != 1:
            raise ValueError("the reference only ever uses num_prob_maps = 1 (Tiling-Shape.py:37)")
        self.debugger = debugger
        self.device = torch.device(device)
        self.complete_graph = complete_graph
        self.network = network
        self.random_network = deepcopy(self.network)      # ml_solver.py:26
        self.num_prob_maps = num_prob_maps

    # ---- resident layout: the greedy loop scores sub-layouts of ONE origin layout (util/algorithms.py:27-31) ---------
    @property
    def supports_node_mask(self):
        """True when ``predict_sub_layout`` can run: the network is a ``tilingnn_b200.TilinGNN`` (subclasses that replace
        ``predict`` with something else, e.g. a CPU checker, do not have one)."""
        return hasattr(getattr(self, "network", None), "set_node_mask")

    def _make_resident(self, layout):
        """Upload ``layout`` once and build its device structures; later calls with the same object are free."""
        if self._is_resident(layout):
            return self._resident[1]
        x, ai, af, ci, _ = to_torch_tensor(self.device, layout.node_feature, layout.align_edge_index,
                                           layout.align_edge_features, layout.collide_edge_index)
        self.network.set_graph(x.shape[0], ai, af, ci)
        self._resident = (layout, x, self.network._native.graph_serial)
        return x

    def _is_resident(self, layout):
        """``layout``'s structures are the ones on the GPU right now (no other graph was set on the network since)."""
        res = getattr(self, "_resident", None)
        return res is not None and res[0] is layout and res[2] == getattr(self.network._native, "graph_serial", None)

    def predict_sub_layout(self, origin_layout, keep):
        """Scores of the sub-layout of ``origin_layout`` induced by the nodes ``keep`` (ascending original indices) -- what
        ``predict(origin_layout.compute_sub_layout(...))`` returns in the reference (util/algorithms.py:27-31,
        tiling/brick_layout.py:248-286), computed WITHOUT re-indexing or re-uploading anything: the origin layout stays
        resident on the GPU and a node mask selects the sub-graph (``tgnn_set_node_mask``)."""
        keep = np.asarray(keep, dtype=np.int64)
        n = origin_layout.node_feature.shape[0]
        if np.size(origin_layout.collide_edge_index) == 0 or np.size(origin_layout.align_edge_index) == 0:
            return np.ones(len(keep), dtype=np.float32)
        x = self._make_resident(origin_layout)
        mask = np.zeros(n, dtype=np.uint8)
        mask[keep] = 1
        _, e_adj, e_col = self.network.set_node_mask(mask)
        if e_col == 0 or e_adj == 0:                                              # ml_solver.py:31-32 on the sub-layout
            return np.ones(len(keep), dtype=np.float32)
        scores = self.network.score(x)
        return scores.detach().cpu().numpy()[keep]

    def predict(self, brick_layout):
        """ml_solver.py:29-49.  Returns ``np.ndarray[N]`` float32."""
        n = brick_layout.node_feature.shape[0]
        # the reference's empty edge set is ``np.array([])`` (len 0); a [2, 0] array means the same here
        if np.size(brick_layout.collide_edge_index) == 0 or np.size(brick_layout.align_edge_index) == 0:
            return np.ones(n, dtype=np.float32)                                   # :31-32
        if self.supports_node_mask and self._is_resident(brick_layout):          # the layout the greedy rounds just ran on
            self.network.set_node_mask(None)
            return self.network.score(self._resident[1]).detach().cpu().numpy()
        x, ai, af, ci, _ = to_torch_tensor(self.device, brick_layout.node_feature, brick_layout.align_edge_index,
                                           brick_layout.align_edge_features, brick_layout.collide_edge_index)
        predictions, *_ = self.network(x=x, adj_e_index=ai, adj_e_features=af, col_e_idx=ci, col_e_features=None)
        # get_best_prob_map (:46,133-136) is argsort over num_prob_maps = 1 losses: always column 0
        return predictions[:, 0].detach().cpu().numpy()

    def solve(self, brick_layout, rng=None, sub_layout="mask"):
        """ml_solver.py:59-67: greedy assembly, then one more scoring pass of the full layout.  Returns
        ``(output_layout, score)``; the layout copy carries ``predict``, ``predict_order``, ``predict_probs``."""
        from . import greedy
        res = greedy.solve_by_probablistic_greedy(self, brick_layout, rng=rng, complete_graph=self.complete_graph,
                                                  sub_layout=sub_layout)
        output_layout = copy(brick_layout)          # the reference deep-copies; nothing mutates the arrays afterwards
        output_layout.predict_order = res.order
        output_layout.predict = res.selection
        output_layout.predict_probs = self.predict(brick_layout)
        output_layout.greedy_rounds = res.rounds
        return output_layout, res.score

    def get_unsupervised_losses_from_layout(self, brick_layout, probs):
        """ml_solver.py:51-57."""
        from . import greedy
        return greedy.calculate_unsupervised_loss(probs, brick_layout.node_feature, brick_layout.collide_edge_index,
                                                  brick_layout.align_edge_index, brick_layout.align_edge_features)[2]

    def get_predict_probs(self, brick_layout):
        """ml_solver.py:69-81."""
        x, ai, af, ci, _ = to_torch_tensor(self.device, brick_layout.node_feature, brick_layout.align_edge_index,
                                           brick_layout.align_edge_features, brick_layout.collide_edge_index)
        return get_network_prediction(self.network, x, ai, af, ci)

    def load_saved_network(self, net_path):
        """ml_solver.py:129-131 -- including the ``.train()`` that makes inference use batch statistics."""
        self.network.load_state_dict(torch.load(net_path, map_location=self.device, weights_only=True))
        self.network.train()
