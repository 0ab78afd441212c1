"""tilingnn_b200 -- B200-native (sm_100a) implementation of TilinGNN's per-node scoring forward pass.

Public surface (mirrors the reference's for this path):
    TilinGNN        graph_networks/networks/TilinGNN.py     (forward / state_dict compatible)
    ML_Solver       solver/ml_solver/ml_solver.py           (predict, load_saved_network)
    get_network_prediction   graph_networks/network_utils.py
"""
from .network import TilinGNN
from .ml_solver import ML_Solver, get_network_prediction, to_torch_tensor

__all__ = ["TilinGNN", "ML_Solver", "get_network_prediction", "to_torch_tensor"]
