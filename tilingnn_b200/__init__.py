"""tilingnn_b200 -- B200-native (sm_100a) implementation of TilinGNN's per-node scoring forward pass.

Public surface (mirrors the reference's for this path):
    TilinGNN        graph_networks/networks/TilinGNN.py     (forward / state_dict compatible)
    ML_Solver       solver/ml_solver/ml_solver.py           (predict, load_saved_network)
    get_network_prediction   graph_networks/network_utils.py
    greedy          util/algorithms.py (probabilistic greedy assembly), brick_layout.compute_sub_layout, losses.py
    tiling_shape    Tiling-Shape.py driver (no plotting)
    tile_graph_io   shapely-free readers of the complete-graph pickles / silhouettes, layout cropping
    ScoreStream     a sequence of host-resident layouts scored with the H2D copy of the next one overlapped (streaming.py)
"""
from .network import TilinGNN
from .ml_solver import ML_Solver, get_network_prediction, to_torch_tensor
from .streaming import ScoreStream, score_stream
from . import greedy, tile_graph_io

__all__ = ["TilinGNN", "ML_Solver", "get_network_prediction", "to_torch_tensor", "ScoreStream", "score_stream"]
