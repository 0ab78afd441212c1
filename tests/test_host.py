"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the
TilinGNN-compatible module has the reference's state_dict layout, loud failure without a GPU,
readers for the reference's data formats, the synthetic generator."""
import copy
import os
import re

import numpy as np
import pytest
import torch

from _util import GOLDEN, load_ckpt, load_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "tgnn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tgnn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    from tilingnn_b200 import _lib
    for name in declared:
        assert hasattr(built_lib, name), f"libtgnn.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table and include/tgnn.h disagree"
    assert built_lib.tgnn_abi_version() == _lib.ABI_VERSION


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(built_lib):
    import ctypes as C
    from tilingnn_b200 import _lib
    cfg = _lib.tgnn_cfg(3, 19, 32, 6, 0, 0)
    h = C.c_void_p()
    assert built_lib.tgnn_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CPU fallback" in built_lib.tgnn_last_error(None)


def test_state_dict_layout_equals_reference_checkpoint():
    from tilingnn_b200 import TilinGNN
    ck = load_ckpt()
    net = TilinGNN(adj_edge_features_dim=19, network_depth=20, network_width=32, node_features_dim=3)
    sd = net.state_dict()
    assert len(sd) == 664
    assert set(sd.keys()) == set(ck.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ck[k].shape), k
    missing, unexpected = net.load_state_dict(ck, strict=True)
    assert not missing and not unexpected
    # the aliased keys are the same storage, as in the reference (edge_conv.py:17-18)
    a = net.state_dict()["brch_1_graph_conv_layers.3.mlp.mlp.2.linear.weight"]
    b = net.state_dict()["brch_1_graph_conv_layers.3.nnConv.nn.mlp.2.linear.weight"]
    assert a.data_ptr() == b.data_ptr()
    # reference key ORDER: compare against the oracle's ordered listing
    from oracle import tilingnn_oracle as orc
    assert list(sd.keys()) == list(orc.reference_param_shapes(3, 19, 20).keys())


def test_module_protocol_train_eval_deepcopy_and_cpu_forward_raises():
    from tilingnn_b200 import TilinGNN
    net = TilinGNN(19, 2, 32, node_features_dim=3)
    assert net.training
    net.eval()
    assert not net.training
    net2 = copy.deepcopy(net)
    assert net2._native is not net._native and net2._native.h is None
    x = torch.zeros(4, 3)
    ei = torch.zeros(2, 3, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(x=x, adj_e_index=ei, adj_e_features=torch.zeros(3, 19), col_e_idx=ei)
    with pytest.raises(ValueError):
        TilinGNN(19, 2, 64, node_features_dim=3)


def test_ml_solver_early_out_and_signature():
    from tilingnn_b200 import ML_Solver, TilinGNN

    class L:
        node_feature = np.zeros((7, 3))
        align_edge_index = np.array([])
        align_edge_features = np.array([])
        collide_edge_index = np.zeros((2, 4), dtype=np.int64)
        collide_edge_features = np.zeros((4, 19))
    s = ML_Solver(None, "cpu", None, TilinGNN(19, 2, 32, node_features_dim=3), 1)
    out = s.predict(L())
    assert out.dtype == np.float32 and out.shape == (7,) and (out == 1).all()
    with pytest.raises(ValueError):
        ML_Solver(None, "cpu", None, TilinGNN(19, 2, 32, node_features_dim=3), 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tilingnn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src.replace("/root/reference/", "REFDOC/") or f.endswith(".py")
    # outside the package only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may touch oracle/
    for f in os.listdir(os.path.join(ROOT, "scripts")):
        if f.endswith((".py", ".sh")):
            assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(ROOT, "scripts", f)).read(), flags=re.M), f


def test_synthetic_generator_properties():
    from tilingnn_b200 import synthetic as syn
    n = 2500
    x, ai, af, ci = syn.lattice_graph(n, 8, 8, seed=3)
    assert x.shape == (n, 3) and ai.dtype == torch.int64 and af.dtype == torch.float32
    ea = set(map(tuple, ai.t().tolist()))
    ec = set(map(tuple, ci.t().tolist()))
    assert len(ea) == ai.shape[1] and len(ec) == ci.shape[1]            # no duplicates
    assert all((b, a) in ea for a, b in ea) and all((b, a) in ec for a, b in ec)   # symmetric
    assert all(a != b for a, b in ea | ec) and not (ea & ec)
    assert 7.5 < ai.shape[1] / n <= 8 and 7.0 < ci.shape[1] / n <= 8
    assert len(torch.unique(af, dim=0)) <= 51
    # feature symmetry feat(u,v) == feat(v,u), like the shipped graphs
    lut = {(int(a), int(b)): af[i] for i, (a, b) in enumerate(ai.t().tolist())}
    for (a, b), f in list(lut.items())[:500]:
        assert torch.equal(f, lut[(b, a)])
    # a destination range generated on its own equals the slice of the whole graph
    x2, ai2, af2, ci2 = syn.lattice_graph(n, 8, 8, seed=3, lo=1000, hi=1700)
    keep = (ai[1] >= 1000) & (ai[1] < 1700)
    assert sorted(map(tuple, ai2.t().tolist())) == sorted(map(tuple, ai[:, keep].t().tolist()))
    assert torch.equal(x2, x[1000:1700])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference data only exists in the build container")
def test_readers_reproduce_the_golden_crop():
    from tilingnn_b200 import tile_graph_io as tio
    z, x, ai, af, ci = load_graph("c1_heart.npz")
    g = tio.load_complete_graph("/root/reference/data/30-60-90/complete_graph_ring9.pkl", tile_type_count=2)
    assert g.num_nodes == 3719 and len(g.adj_edges) == 54920 and len(g.colli_edges) == 122904
    ext, ints = tio.load_polygons("/root/reference/silhouette/heart.txt")
    crops = tio.crop_multiple_layouts_from_contour(ext, ints, g, 0, 30, 1, [0, 0.5], [0.5])
    sizes = [[c.node_feature.shape[0], c.align_edge_index.shape[1], c.collide_edge_index.shape[1]] for c in crops]
    assert sizes == z["crop_sizes"].tolist()
    assert np.array_equal(crops[0].tiles, z["tiles"])
    assert np.allclose(crops[0].node_feature, z["x"])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference data only exists in the build container")
def test_tiling_shape_driver_on_the_reference_files(tmp_path):
    """Tiling-Shape.py end to end on the reference's own files (pickle, checkpoint, silhouette) with the fp32 oracle network
    standing in for the CUDA one: four bunny layouts, valid maximal selections, result pickles written."""
    import pickle
    from oracle import tilingnn_oracle as orc
    from tilingnn_b200 import ML_Solver, tiling_shape

    class OracleSolver(ML_Solver):
        def __init__(self, g, state):
            self.complete_graph, self.params = g, state

        def predict(self, lay):
            n = lay.node_feature.shape[0]
            if np.size(lay.collide_edge_index) == 0 or np.size(lay.align_edge_index) == 0:
                return np.ones(n, dtype=np.float32)
            t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
            s = orc.forward(self.params, t(lay.node_feature, torch.float32), t(lay.align_edge_index, torch.long),
                            t(lay.align_edge_features, torch.float32), t(lay.collide_edge_index, torch.long), depth=20,
                            bn_mode="train", dtype=torch.float32)
            return s[:, 0].numpy()
    ref = "/root/reference"
    sols = tiling_shape.tiling_a_region(f"{ref}/data/30-60-90+equilateral/complete_graph_ring9.pkl",
                                        f"{ref}/pre-trained_models/30-60-90+equilateral.pth", f"{ref}/silhouette/bunny.txt",
                                        out_dir=str(tmp_path), verbose=False, solver_factory=OracleSolver)
    assert [s.node_feature.shape[0] for s, _ in sols] == [604, 562, 591, 565]
    for lay, score in sols:
        sel = lay.predict.astype(bool)
        ci = lay.collide_edge_index
        assert not (sel[ci[0]] & sel[ci[1]]).any() and 0 < score <= 1.02
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 4
    d = pickle.load(open(os.path.join(tmp_path, files[0]), "rb"))
    assert set(d) == {"tiles", "predict", "predict_order", "predict_probs", "score"}


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference data only exists in the build container")
def test_batched_clip_equals_per_tile_clip():
    """the vectorised Sutherland-Hodgman of crop_from_contour against the per-tile implementation (areas and kept tiles)"""
    from tilingnn_b200 import tile_graph_io as tio
    g = tio.load_complete_graph("/root/reference/data/45-45-90+rectangle/complete_graph_ring9.pkl", tile_type_count=2)
    for name, margin, ang, dx in (("heart.txt", 0.5, 0.0, 0.0), ("bunny.txt", 0.7, 15.0, 0.3)):
        ext, ints = tio.load_polygons(f"/root/reference/silhouette/{name}")
        _, e2, i2 = tio.shape_transform(g, ext, ints, margin, ang, dx, dx)
        idx = list(range(0, g.num_nodes, 7))
        rings = [g.tile_rings[i] for i in idx]
        batch = tio.intersection_areas_with_convex(e2, i2, rings)
        single = np.asarray([tio.intersection_area_with_convex(e2, i2, r) for r in rings])
        assert np.allclose(batch, single, rtol=0, atol=1e-12)
        assert (np.abs(batch - g.tile_areas[idx]) < 1e-6).any() and (batch == 0).any()      # contained and disjoint tiles both occur


def test_graph_unpickler_rejects_globals_outside_the_allow_list():
    """complete_graph_ring*.pkl are external data files: the stub unpickler must not import arbitrary callables."""
    import io
    import pickle
    from tilingnn_b200 import tile_graph_io as tio

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))
    with pytest.raises(pickle.UnpicklingError, match="allow-list"):
        tio._GraphUnpickler(io.BytesIO(pickle.dumps({"tiles": [Evil()]}))).load()
    ok = tio._GraphUnpickler(io.BytesIO(pickle.dumps({"a": np.arange(3), "b": [1.0, (2, 3)]}))).load()
    assert ok["a"].tolist() == [0, 1, 2]


def test_even_bounds_rejects_empty_ranges_on_every_rank():
    from tilingnn_b200 import shard
    assert shard.even_bounds(1000, 4) == [0, 256, 512, 768, 1000]
    with pytest.raises(ValueError, match="non-empty"):
        shard.even_bounds(128, 4)          # 64-aligned ranges would leave ranks 2 and 3 empty
