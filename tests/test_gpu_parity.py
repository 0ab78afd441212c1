"""Parity of the CUDA path (through the C ABI) with the oracle and the committed golden vectors.
Run on the B200 box:  python -m pytest tests -m gpu -x -q

Tolerances (fp32 arithmetic):
  tier 1  per-layer tensors vs fp64 reference on identical inputs: <= 1e-4 relative to max-abs
  tier 2  end-to-end, well-conditioned nets (eval-BN on the shipped checkpoint; synthetic
          conditioned weights in train-BN): <= 1e-4 absolute on the [0,1] scores
  tier 3  end-to-end train-BN on the shipped checkpoint (ill-conditioned: the reference's own fp32
          run is 1e-3..7e-2 away from fp64): ours must be within 1e-3 absolute of fp64 (measured
          2.0e-4 / 4.1e-4) AND no further from it than the reference's fp32 run is, and rank the nodes
          the same way (top-100 overlap >= 98; measured 100 / 99).
"""
import numpy as np
import pytest
import torch

from oracle import tilingnn_oracle as orc
from _util import load_ckpt, load_graph, syn_small_params

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def make_net(params, d_x, d_e, depth, dev, mode="train"):
    from tilingnn_b200 import TilinGNN
    net = TilinGNN(d_e, depth, 32, node_features_dim=d_x)
    net.load_state_dict(params, strict=True)
    net = net.to(dev)
    return net.train() if mode == "train" else net.eval()


def run(net, x, ai, af, ci, dev):
    s, feats = net(x=x.to(dev), adj_e_index=ai.to(dev), adj_e_features=af.to(dev), col_e_idx=ci.to(dev))
    torch.cuda.synchronize()
    assert s.shape == (x.shape[0], 1) and s.dtype == torch.float32
    return s[:, 0].double().cpu().numpy()


def relerr(a, b):
    return np.abs(a - b).max() / max(1e-30, np.abs(b).max())


# ---------------------------------------------------------------------------------------------- #
def test_graph_structures_are_a_faithful_reencoding(dev):
    """every adjacency edge sits in exactly one slot of a chunk of its type; destinations are distinct
    inside each 8-slot group; collision CSR = edges sorted by destination without self loops."""
    z, x, ai, af, ci = load_graph("syn_small.npz")
    p, depth = syn_small_params(z)
    net = make_net(p, x.shape[1], af.shape[1], depth, dev)
    net.set_graph(x.shape[0], ai.to(dev), af.to(dev), ci.to(dev))
    g = net.debug_graph()
    info = net.info()
    n = x.shape[0]
    assert info["e_adj"] == ai.shape[1]
    rows = g["type_rows"].numpy()
    assert info["n_edge_types"] == len(np.unique(af.numpy(), axis=0)) == rows.shape[0]
    cptr, ctype, csrc, cdst = g["cptr"].numpy(), g["ctype"].numpy(), g["csrc"].numpy(), g["cdst"].numpy()
    assert cptr[0] == 0 and cptr[-1] == len(ctype) and (np.diff(cptr) >= 0).all()
    got = []
    for t in range(len(cptr) - 1):
        for c in range(cptr[t], cptr[t + 1]):
            for grp in range(2):
                sl = slice(c * 16 + grp * 8, c * 16 + grp * 8 + 8)
                live = csrc[sl] >= 0
                d = cdst[sl][live]
                assert len(set(d.tolist())) == len(d), "duplicate destination inside a group"
                for s_, d_ in zip(csrc[sl][live], d):
                    got.append((int(s_), t * info["tile_rows"] + int(d_), tuple(rows[ctype[c]])))
    want = [(int(a), int(b), tuple(f)) for (a, b), f in zip(ai.t().tolist(), af.numpy())]
    assert sorted(got) == sorted(want)
    deg = np.bincount(ai[1].numpy(), minlength=n)
    assert np.allclose(g["inv_deg"].numpy(), 1.0 / np.maximum(deg, 1))
    keep = ci[0] != ci[1]
    cs, cd = ci[0][keep].numpy(), ci[1][keep].numpy()
    order = np.argsort(cd, kind="stable")
    assert np.array_equal(g["col_src"].numpy(), cs[order])
    assert np.array_equal(np.diff(g["col_ptr"].numpy()), np.bincount(cd, minlength=n))


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_ragged_synthetic_vs_reference_golden(dev, mode):
    z, x, ai, af, ci = load_graph("syn_small.npz")
    p, depth = syn_small_params(z)
    net = make_net(p, x.shape[1], af.shape[1], depth, dev, mode)
    s = run(net, x, ai, af, ci, dev)
    assert np.abs(s - z[f"ref_{mode}_f64"]).max() <= TOL


def test_per_layer_tensors_vs_reference_golden(dev):
    """tier 1 on the shipped checkpoint / heart crop: h0, then layer outputs g1, g2 after 1, 2, 6 layers."""
    z, x, ai, af, ci = load_graph("c1_heart.npz")
    net = make_net(load_ckpt(), 3, 19, 20, dev)
    for i in (0, 1, 5):
        net.debug_set_stop_layer(i)
        net(x=x.to(dev), adj_e_index=ai.to(dev), adj_e_features=af.to(dev), col_e_idx=ci.to(dev))
        if i == 0:
            assert relerr(net.debug_read("mid_0").double().cpu().numpy(), z["train_h0"]) <= TOL
        g1 = net.debug_read("g1").double().cpu().numpy()
        g2 = net.debug_read("g2").double().cpu().numpy()
        e1, e2 = relerr(g1, z[f"train_g1_{i}"]), relerr(g2, z[f"train_g2_{i}"])
        print(f"layer {i}: g1 rel err {e1:.2e}  g2 rel err {e2:.2e}")
        lim = TOL if i < 2 else 30 * TOL          # error compounds through ill-conditioned BN
        assert e1 <= lim and e2 <= lim
    net.debug_set_stop_layer(-1)


@pytest.mark.parametrize("graph", ["c1_heart.npz", "c1_complete.npz"])
def test_config1_eval_bn_end_to_end(dev, graph):
    """tier 2: shipped checkpoint, eval-BN."""
    z, x, ai, af, ci = load_graph(graph)
    net = make_net(load_ckpt(), 3, 19, 20, dev, "eval")
    s = run(net, x, ai, af, ci, dev)
    ours = np.abs(s - z["ref_eval_f64"]).max()
    ref32 = np.abs(z["ref_eval_f32"] - z["ref_eval_f64"]).max()
    print(f"{graph} eval-BN: ours {ours:.2e}  reference-fp32 {ref32:.2e}")
    # the reference's own fp32 run is 2.7e-4 from fp64 here (collapsed running_var in the collision branch: gains of
    # ~316 on fp32 rounding); ours must not be worse than it, with a hard cap of 3e-4
    assert ours <= max(TOL, 1.1 * ref32) and ours <= 3.5e-4


@pytest.mark.parametrize("graph", ["c1_heart.npz", "c1_complete.npz"])
def test_config1_train_bn_end_to_end(dev, graph):
    """tier 3: shipped checkpoint, train-BN (the reference's behaviour)."""
    z, x, ai, af, ci = load_graph(graph)
    net = make_net(load_ckpt(), 3, 19, 20, dev, "train")
    s = run(net, x, ai, af, ci, dev)
    gold = z["ref_train_f64"]
    ours, ref32 = np.abs(s - gold).max(), np.abs(z["ref_train_f32"] - gold).max()
    top = lambda v: set(np.argsort(-v)[:100].tolist())
    overlap, ref_overlap = len(top(s) & top(gold)), len(top(z["ref_train_f32"]) & top(gold))
    print(f"{graph} train-BN: ours max {ours:.2e} mean {np.abs(s - gold).mean():.2e} | reference-fp32 max {ref32:.2e} "
          f"mean {np.abs(z['ref_train_f32'] - gold).mean():.2e} | top-100 overlap ours {overlap} ref {ref_overlap}")
    assert ours <= 1e-3 and ours <= max(TOL, ref32)
    assert overlap >= 98


@pytest.mark.parametrize("n,deg", [(10000, 8), (4000, 32)])
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_synthetic_configs_vs_oracle(dev, n, deg, mode):
    """config 2 (10k nodes, deg 8+8, 6 layers) and a deg-32 case, conditioned weights."""
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(n, deg, deg, seed=0)
    p = orc.make_params(3, 19, 6, seed=0)
    if mode == "eval":      # running statistics = this graph's batch statistics, else the net saturates to 0
        p = orc.calibrate_running_stats(p, x, ai, af, ci, depth=6)
    gold = orc.forward(p, x, ai, af, ci, depth=6, bn_mode=mode, dtype=torch.float64)[:, 0].numpy()
    assert gold.std() > 0.05
    net = make_net(p, 3, 19, 6, dev, mode)
    s = run(net, x, ai, af, ci, dev)
    err = np.abs(s - gold).max()
    print(f"N={n} deg={deg} {mode}: max err {err:.2e}")
    assert err <= TOL


@pytest.mark.parametrize("kernel", ["chunk", "s", "h", "t", "z", "z32", "x"])
def test_all_adjacency_kernels(dev, kernel, monkeypatch):
    """The fp16-split edge-chunk kernel, its 3xTF32 twin, the tcgen05 S kernel, the tcgen05 edge-block kernel and the
    windowed tcgen05 kernel (A operand in tensor memory) are
    selected per graph by size / a cost model; force each (TGNN_CONV) on the shipped checkpoint (20 edge types) and on a
    synthetic graph (51 types)."""
    monkeypatch.setenv("TGNN_CONV", kernel)
    from tilingnn_b200 import synthetic as syn
    z, x, ai, af, ci = load_graph("c1_complete.npz")
    net = make_net(load_ckpt(), 3, 19, 20, dev, "train")
    s = run(net, x, ai, af, ci, dev)
    gold = z["ref_train_f64"]
    assert np.abs(s - gold).max() <= 1e-3          # tier 3 (reference fp32: 6.6e-2)
    x, ai, af, ci = syn.lattice_graph(6000, 32, 32, seed=3)
    p = orc.make_params(3, 19, 4, seed=3)
    gold = orc.forward(p, x, ai, af, ci, depth=4, dtype=torch.float64)[:, 0].numpy()
    net = make_net(p, 3, 19, 4, dev)
    err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
    print(f"TGNN_CONV={kernel}: max err {err:.2e}")
    assert err <= TOL
    info = net.info()
    assert info["conv_kernel"] == {"chunk": 0, "s": 1, "h": 2, "t": 3, "z": 4, "z32": 4, "x": 2}[kernel] and info["range_fallback_layers"] == 0


@pytest.mark.parametrize("variant", ["z", "z32"])
def test_conv_z_multi_edges_isolated_rows_and_mask(dev, variant, monkeypatch):
    """k_conv_z builds one A row per (destination, edge type): duplicate edges and several same-type in-edges are summed
    in fp32 by the gather thread, destinations without in-edges give zero rows (mean -> 0, root term only), self loops
    are ordinary adjacency edges.  Lattice graph with all of these added, against the fp64 oracle."""
    monkeypatch.setenv("TGNN_CONV", variant)
    from tilingnn_b200 import synthetic as syn
    n = 5000
    x, ai, af, ci = syn.lattice_graph(n, 16, 16, seed=5)
    g = torch.Generator().manual_seed(7)
    keep = (ai[1] % 37 != 3)                                    # destinations 3, 40, 77, ... lose all in-edges
    ai, af = ai[:, keep], af[keep]
    dup = torch.randint(0, ai.shape[1], (4000,), generator=g)   # duplicate edges (same feature row -> same type)
    loops = torch.arange(0, n, 11)
    ai = torch.cat([ai, ai[:, dup], torch.stack([loops, loops])], 1)
    af = torch.cat([af, af[dup], af[:loops.numel()]], 0)
    p = orc.make_params(3, 19, 4, seed=5)
    for mode in ("train", "eval"):
        q = orc.calibrate_running_stats(p, x, ai, af, ci, depth=4) if mode == "eval" else p
        gold = orc.forward(q, x, ai, af, ci, depth=4, bn_mode=mode, dtype=torch.float64)[:, 0].numpy()
        net = make_net(q, 3, 19, 4, dev, mode)
        err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
        net.check_errors()
        print(f"k_conv_z multi-edge lattice {mode}: max err {err:.2e}")
        assert net.info()["conv_kernel"] == 4 and err <= TOL


def test_gin_mlp_on_3xtf32(dev, monkeypatch):
    """the GIN MLP's layers 2-3 run on fp16-split operands by default; TGNN_GIN=tf32 keeps all three on 3xTF32 (the
    variant used when a GIN weight is outside the fp16 range)."""
    monkeypatch.setenv("TGNN_GIN", "tf32")
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(6000, 32, 32, seed=3)
    p = orc.make_params(3, 19, 4, seed=3)
    gold = orc.forward(p, x, ai, af, ci, depth=4, dtype=torch.float64)[:, 0].numpy()
    err = np.abs(run(make_net(p, 3, 19, 4, dev), x, ai, af, ci, dev) - gold).max()
    print(f"TGNN_GIN=tf32: max err {err:.2e}")
    assert err <= TOL


@pytest.mark.parametrize("kernel", ["h", "t", "z"])
def test_fp16_range_guard_hands_layers_to_the_tf32_kernel(dev, kernel, monkeypatch):
    """k_conv_h / k_conv_t / k_conv_z<fp16> work on fp16-split operands; activations beyond +-60000 (here: a BatchNorm gain of 1e6 in
    layer 0) and root weights beyond it (layer 2) must raise the range flags so the 3xTF32 arithmetic (k_conv_h) / the
    fp32 stand-by kernel (k_conv_t) / the tf32 variant (k_conv_z) takes those layers."""
    monkeypatch.setenv("TGNN_CONV", kernel)
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(3000, 8, 8, seed=4)
    p = dict(orc.make_params(3, 19, 4, seed=4))
    p["brch_1_graph_conv_layers.0.batch_norm.weight"] = p["brch_1_graph_conv_layers.0.batch_norm.weight"] * 1e6
    p["brch_1_graph_conv_layers.2.nnConv.root"] = p["brch_1_graph_conv_layers.2.nnConv.root"] * 1e6
    gold = orc.forward(p, x, ai, af, ci, depth=4, dtype=torch.float64)[:, 0].numpy()
    net = make_net(p, 3, 19, 4, dev)
    err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
    info = net.info()
    print(f"range guard: max err {err:.2e}, fallback layers {info['range_fallback_layers']}")
    assert info["conv_kernel"] == {"h": 2, "t": 3, "z": 4}[kernel] and info["range_fallback_layers"] >= 2
    assert err <= TOL


def test_random_graph_and_many_edge_types(dev):
    """uniform random sources (duplicates, self loops) and continuous features (one type per edge)."""
    from tilingnn_b200 import synthetic as syn
    p = orc.make_params(3, 19, 3, seed=1)
    x, ai, af, ci = syn.random_graph(1500, 6, 10, seed=5)
    gold = orc.forward(p, x, ai, af, ci, depth=3, dtype=torch.float64)[:, 0].numpy()
    net = make_net(p, 3, 19, 3, dev)
    assert np.abs(run(net, x, ai, af, ci, dev) - gold).max() <= TOL
    x, ai, af, ci = syn.lattice_graph(900, 8, 8, seed=2, continuous_features=True)
    gold = orc.forward(p, x, ai, af, ci, depth=3, dtype=torch.float64)[:, 0].numpy()
    assert np.abs(run(net, x, ai, af, ci, dev) - gold).max() <= TOL
    assert net.info()["n_edge_types"] > 1000


def test_more_than_65535_edge_types_streams_the_weight_tables(dev):
    """continuous edge features on ~158k edges (79k distinct rows: the generator is symmetric); the per-type weight tables (12 KB each)
    are then built one layer at a time instead of for all layers at once."""
    from tilingnn_b200 import synthetic as syn
    p = orc.make_params(3, 19, 3, seed=1)
    x, ai, af, ci = syn.lattice_graph(20000, 8, 8, seed=6, continuous_features=True)
    gold = orc.forward(p, x, ai, af, ci, depth=3, dtype=torch.float64)[:, 0].numpy()
    net = make_net(p, 3, 19, 3, dev)
    err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
    k = net.info()["n_edge_types"]
    print(f"{k} edge types: max err {err:.2e}")
    assert k > 65535 and err <= TOL


def test_degenerate_inputs(dev):
    """no adjacency edges / no collision edges / a single node; bad indices raise."""
    from tilingnn_b200 import synthetic as syn
    p = orc.make_params(3, 19, 2, seed=1)
    net = make_net(p, 3, 19, 2, dev)
    x, ai, af, ci = syn.lattice_graph(300, 8, 8, seed=2)
    e0 = torch.zeros(2, 0, dtype=torch.int64)
    gold = orc.forward(p, x, e0, af[:0], ci, depth=2, dtype=torch.float64)[:, 0].numpy()
    assert np.abs(run(net, x, e0, af[:0], ci, dev) - gold).max() <= TOL
    gold = orc.forward(p, x, ai, af, e0, depth=2, dtype=torch.float64)[:, 0].numpy()
    assert np.abs(run(net, x, ai, af, e0, dev) - gold).max() <= TOL
    bad = ai.clone()
    bad[0, 5] = 300
    with pytest.raises(RuntimeError, match="out of range"):
        run(net, x, bad, af, ci, dev)
    with pytest.raises(ValueError):
        run(net, x, ai, af[:, :5], ci, dev)


def test_run_to_run_determinism_and_input_immutability(dev):
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(20000, 8, 8, seed=0)
    p = orc.make_params(3, 19, 6, seed=0)
    net = make_net(p, 3, 19, 6, dev)
    xd, aid, afd, cid = x.to(dev), ai.to(dev), af.to(dev), ci.to(dev)
    keep = [t.clone() for t in (xd, aid, afd, cid)]
    a, feats = net(x=xd, adj_e_index=aid, adj_e_features=afd, col_e_idx=cid)
    a = a.clone()
    assert feats is afd                                           # the reference returns adj_e_features untouched
    net.set_graph(x.shape[0], aid, afd, cid)                       # rebuild everything
    b = net.score(xd).clone()
    torch.cuda.synchronize()
    assert torch.equal(a[:, 0], b), "two runs on the same input must be bit-identical"
    for t, k in zip((xd, aid, afd, cid), keep):
        assert torch.equal(t, k)


def test_edge_order_invariance_at_scale(dev):
    """size-independent property: permuting the edge lists changes nothing beyond fp32 summation
    order (checked on a 200k-node graph the oracle would not finish in seconds)."""
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(200000, 32, 32, seed=0, device=dev)
    p = orc.make_params(3, 19, 6, seed=0)
    net = make_net(p, 3, 19, 6, dev)
    a = net(x=x, adj_e_index=ai, adj_e_features=af, col_e_idx=ci)[0].clone()
    g = torch.Generator(device="cpu").manual_seed(0)
    pa = torch.randperm(ai.shape[1], generator=g).to(dev)
    pc = torch.randperm(ci.shape[1], generator=g).to(dev)
    b = net(x=x, adj_e_index=ai[:, pa].contiguous(), adj_e_features=af[pa].contiguous(), col_e_idx=ci[:, pc].contiguous())[0]
    torch.cuda.synchronize()
    assert torch.isfinite(a).all() and (a >= 0).all() and (a <= 1).all()
    assert (a - b).abs().max().item() <= TOL
    info = net.info()
    assert info["n_edge_types"] <= 51 and info["adj_slots"] >= info["e_adj"]


def test_ml_solver_predict_matches_oracle(dev):
    """the reference-facing call: numpy layout in, numpy scores out (ml_solver.py:29-49)."""
    from tilingnn_b200 import ML_Solver
    z, x, ai, af, ci = load_graph("c1_heart.npz")

    class Layout:
        node_feature = x.double().numpy()
        align_edge_index = ai.numpy()
        align_edge_features = af.double().numpy()
        collide_edge_index = ci.numpy()
        collide_edge_features = np.zeros((ci.shape[1], 19))
    net = make_net(load_ckpt(), 3, 19, 20, dev, "eval")
    solver = ML_Solver(None, dev, None, net, 1)
    out = solver.predict(Layout())
    assert out.dtype == np.float32 and out.shape == (x.shape[0],)
    assert np.abs(out - z["ref_eval_f64"]).max() <= 3.5e-4         # tier 2 cap (reference fp32: 2.8e-4)


def test_get_network_prediction_and_get_predict_probs(dev, capsys):
    """a7: ``get_network_prediction`` (graph_networks/network_utils.py:4-21) returns ``probs`` only, keyword-calls the
    network, prints the traceback and re-raises on failure; ``ML_Solver.get_predict_probs`` (ml_solver.py:69-81) is the
    same call from a layout.  Both must give exactly what ``network.forward`` gives."""
    from tilingnn_b200 import ML_Solver
    from tilingnn_b200.ml_solver import get_network_prediction
    z, x, ai, af, ci = load_graph("c1_heart.npz")
    net = make_net(load_ckpt(), 3, 19, 20, dev, "train")
    xd, aid, afd, cid = x.to(dev), ai.to(dev), af.to(dev), ci.to(dev)
    direct, feats = net(x=xd, adj_e_index=aid, adj_e_features=afd, col_e_idx=cid)
    direct = direct.clone()
    probs = get_network_prediction(net, xd, aid, afd, cid, None)
    assert isinstance(probs, torch.Tensor) and probs.shape == (x.shape[0], 1) and probs.device.type == "cuda"
    assert torch.equal(probs, direct)
    gold = z["ref_train_f64"]
    assert np.abs(probs[:, 0].double().cpu().numpy() - gold).max() <= 1e-3

    class Layout:
        node_feature = x.double().numpy()
        align_edge_index = ai.numpy()
        align_edge_features = af.double().numpy()
        collide_edge_index = ci.numpy()
        collide_edge_features = np.zeros((ci.shape[1], 19))
    solver = ML_Solver(None, dev, None, net, 1)
    p2 = solver.get_predict_probs(Layout())
    assert p2.shape == (x.shape[0], 1) and torch.equal(p2, direct)
    assert np.array_equal(solver.predict(Layout()), direct[:, 0].cpu().numpy())
    # failure contract: traceback printed, exception re-raised (network_utils.py:16-19)
    with pytest.raises(RuntimeError):
        get_network_prediction(net, xd.cpu(), aid, afd, cid)
    assert "Traceback" in capsys.readouterr().out


def test_gin_staged_window_kernel(dev, monkeypatch):
    """k_gin_w (neighbour rows staged in shared-memory windows by TMA bulk copies) is chosen by itself only for large
    graphs; TGNN_GINW=1 forces it here on the small parity cases: lattices (all tiles get a window), the shipped
    checkpoint's complete graph (ragged degrees), a partial last tile, and a uniformly random graph whose tiles are
    not local and must take the in-kernel 'direct' path.  TGNN_GINW=0 must give the per-lane-gather kernel."""
    from tilingnn_b200 import synthetic as syn
    monkeypatch.setenv("TGNN_GINW", "1")
    for n, deg, depth, seed in ((4000, 32, 4, 3), (10000, 8, 6, 0), (1000 + 37, 8, 3, 5)):
        x, ai, af, ci = syn.lattice_graph(n, deg, deg, seed=seed)
        p = orc.make_params(3, 19, depth, seed=seed)
        gold = orc.forward(p, x, ai, af, ci, depth=depth, dtype=torch.float64)[:, 0].numpy()
        net = make_net(p, 3, 19, depth, dev)
        err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
        info = net.info()
        print(f"k_gin_w lattice N={n} deg={deg}: max err {err:.2e}  window tiles {info['gin_window_tiles']} direct {info['gin_direct_tiles']}")
        assert info["gin_kernel"] == 1 and info["gin_direct_tiles"] == 0 and info["gin_window_tiles"] == (n + 63) // 64
        assert err <= TOL
    z, x, ai, af, ci = load_graph("c1_complete.npz")
    net = make_net(load_ckpt(), 3, 19, 20, dev, "train")
    s = run(net, x, ai, af, ci, dev)
    info = net.info()
    print(f"k_gin_w complete graph: max err {np.abs(s - z['ref_train_f64']).max():.2e}  window tiles {info['gin_window_tiles']} "
          f"direct {info['gin_direct_tiles']}")
    assert info["gin_kernel"] == 1 and np.abs(s - z["ref_train_f64"]).max() <= 1e-3
    net = make_net(load_ckpt(), 3, 19, 20, dev, "eval")
    assert np.abs(run(net, x, ai, af, ci, dev) - z["ref_eval_f64"]).max() <= 3.5e-4
    p = orc.make_params(3, 19, 3, seed=1)
    x, ai, af, ci = syn.random_graph(3000, 6, 10, seed=5)
    gold = orc.forward(p, x, ai, af, ci, depth=3, dtype=torch.float64)[:, 0].numpy()
    net = make_net(p, 3, 19, 3, dev)
    err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
    info = net.info()
    print(f"k_gin_w random graph: max err {err:.2e}  window tiles {info['gin_window_tiles']} direct {info['gin_direct_tiles']}")
    assert info["gin_direct_tiles"] > 0 and err <= TOL
    monkeypatch.setenv("TGNN_GINW", "0")
    net = make_net(p, 3, 19, 3, dev)
    assert np.abs(run(net, x, ai, af, ci, dev) - gold).max() <= TOL and net.info()["gin_kernel"] == 0


def test_edge_block_format_is_a_faithful_reencoding(dev, monkeypatch):
    """T format of the tcgen05 edge-block kernel: every adjacency edge sits in exactly one slot of a block of its type;
    quarter q of a block only holds destinations with dst % 4 == q, no destination twice inside a quarter (that is what
    lets four epilogue warps accumulate into one shared tile without atomics); every super-tile ends with its root blocks."""
    monkeypatch.setenv("TGNN_CONV", "t")
    for name in ("syn_small.npz", "c1_complete.npz"):
        z, x, ai, af, ci = load_graph(name)
        if name == "syn_small.npz":
            p, depth = syn_small_params(z)
        else:
            p, depth = load_ckpt(), 20
        net = make_net(p, x.shape[1], af.shape[1], depth, dev)
        net.set_graph(x.shape[0], ai.to(dev), af.to(dev), ci.to(dev))
        info = net.info()
        g, gt = net.debug_graph(), net.debug_graph_t()
        rows, n, rt, K = g["type_rows"].numpy(), x.shape[0], info["t_rows"], info["n_edge_types"]
        bptr, btype, tsrc, tdst = (gt[k].numpy() for k in ("bptr", "btype", "tsrc", "tdst"))
        assert rt == 256 and bptr[0] == 0 and bptr[-1] == info["t_blocks"] == len(btype)
        got, roots = [], []
        for t in range(len(bptr) - 1):
            assert bptr[t + 1] - bptr[t] >= rt // 128
            for b in range(bptr[t], bptr[t + 1]):
                is_root = btype[b] == K
                assert is_root == (b >= bptr[t + 1] - rt // 128), "root blocks must be the last blocks of a super-tile"
                for q in range(4):
                    sl = slice(b * 128 + 32 * q, b * 128 + 32 * q + 32)
                    live = tsrc[sl] >= 0
                    d = tdst[sl][live]
                    assert (tdst[sl][~live] == 0xFFFF).all()
                    assert (d % 4 == q).all(), "destination class"
                    assert len(set(d.tolist())) == len(d), "duplicate destination inside a quarter"
                    for s_, d_ in zip(tsrc[sl][live], d):
                        if is_root:
                            roots.append((int(s_), t * rt + int(d_)))
                        else:
                            got.append((int(s_), t * rt + int(d_), tuple(rows[btype[b]])))
        want = [(int(a), int(b), tuple(f)) for (a, b), f in zip(ai.t().tolist(), af.numpy())]
        assert sorted(got) == sorted(want)
        assert sorted(roots) == [(i, i) for i in range(n)]


def test_graph_replay_of_repeated_forwards(dev):
    """Forwards on a resident graph: the first runs eagerly, the second is captured into a CUDA graph, later ones replay it
    (one submission instead of ~75 launches for a real layout).  All of them must be bit-identical, follow a new x, and
    survive a change of graph, of parameters and of BatchNorm mode."""
    from tilingnn_b200 import synthetic as syn
    z, x, ai, af, ci = load_graph("c1_heart.npz")
    net = make_net(load_ckpt(), 3, 19, 20, dev, "train")
    xd, aid, afd, cid = x.to(dev), ai.to(dev), af.to(dev), ci.to(dev)
    net.set_graph(x.shape[0], aid, afd, cid)
    outs = [net.score(xd).clone() for _ in range(4)]           # eager, capture + replay, replay, replay
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    assert np.abs(outs[0].double().cpu().numpy() - z["ref_train_f64"]).max() <= 1e-3
    x2 = xd.clone(); x2[:, -1] *= 0.5                          # another input through the SAME captured sequence
    a = net.score(x2).clone()
    net2 = make_net(load_ckpt(), 3, 19, 20, dev, "train")
    net2.set_graph(x.shape[0], aid, afd, cid)
    assert torch.equal(a, net2.score(x2)), "replay must read the caller's x, not the captured one"
    assert not torch.equal(a, outs[0])
    net.eval()                                                  # mode change: new key, eager again, then capture
    e = [net.score(xd).clone() for _ in range(3)]
    assert all(torch.equal(e[0], o) for o in e[1:])
    assert np.abs(e[0].double().cpu().numpy() - z["ref_eval_f64"]).max() <= 3.5e-4
    net.train()
    xs, ais, afs, cis = syn.lattice_graph(3000, 8, 8, seed=0)   # other graph, other parameters on the same handle
    p = orc.make_params(3, 19, 20, seed=0)
    net.load_state_dict(p, strict=True)
    gold = orc.forward(p, xs, ais, afs, cis, depth=20, dtype=torch.float64)[:, 0].numpy()
    net.set_graph(3000, ais.to(dev), afs.to(dev), cis.to(dev))
    for _ in range(3):
        s = net.score(xs.to(dev))
        assert np.abs(s.double().cpu().numpy() - gold).max() <= TOL


@pytest.mark.parametrize("var,val", [("TGNN_GINW", "1"), ("TGNN_CONV", "t"), ("TGNN_CONV", "z"), ("TGNN_CONV", "x")])
def test_persistent_pipelines_over_many_tiles(dev, var, val, monkeypatch):
    """k_gin_w and k_conv_t are persistent kernels whose mbarrier pipelines run for dozens of tiles per CTA at benchmark
    sizes (the small parity cases give every CTA a single tile).  Size-independent property on a 300k-node graph the
    oracle would not finish: the staged-window / tcgen05 kernel must agree with the per-lane-gather / mma.sync kernel."""
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(300000, 16, 16, seed=0, device=dev)
    p = orc.make_params(3, 19, 3, seed=0)
    monkeypatch.setenv("TGNN_GINW", "0"); monkeypatch.setenv("TGNN_CONV", "h")
    base = make_net(p, 3, 19, 3, dev)
    a = base(x=x, adj_e_index=ai, adj_e_features=af, col_e_idx=ci)[0].clone()
    monkeypatch.setenv(var, val)
    net = make_net(p, 3, 19, 3, dev)
    b = net(x=x, adj_e_index=ai, adj_e_features=af, col_e_idx=ci)[0].clone()
    net.check_errors()
    info = net.info()
    assert (info["gin_kernel"] == 1) if var == "TGNN_GINW" else (info["conv_kernel"] == {"t": 3, "z": 4, "x": 5}[val])
    err = (a - b).abs().max().item()
    print(f"{var}={val} vs baseline kernels on 300k nodes: max diff {err:.2e}")
    assert torch.isfinite(b).all() and err <= 2e-5


def test_score_stream_matches_serial_calls(dev):
    """ScoreStream (double-buffered device slots, H2D copy of layout k+1 overlapped with the scoring of layout k) must
    return exactly what one serial call per layout returns -- different graphs of different sizes in one stream."""
    from tilingnn_b200 import ScoreStream, synthetic as syn
    p = orc.make_params(3, 19, 3, seed=2)
    net = make_net(p, 3, 19, 3, dev)
    graphs = [syn.lattice_graph(n, d, d, seed=s) for n, d, s in ((3000, 8, 0), (5000, 16, 1), (3000, 8, 2), (777, 4, 3), (5000, 16, 1))]
    serial = [run(net, *g, dev) for g in graphs]
    pinned = [[t.pin_memory() for t in g] for g in graphs]
    outs = [torch.empty(g[0].shape[0], dtype=torch.float32).pin_memory() for g in graphs]
    ScoreStream(net)(pinned, outs)
    torch.cuda.synchronize()
    for a, b in zip(serial, outs):
        assert np.array_equal(a, b.numpy())
    gold = orc.forward(p, *graphs[1], depth=3, dtype=torch.float64)[:, 0].numpy()
    assert np.abs(outs[1].numpy() - gold).max() <= TOL and np.array_equal(outs[1].numpy(), outs[4].numpy())


@pytest.mark.parametrize("variant", ["fp16", "tf32"])
def test_final_mlp_variants_and_range_guard(dev, variant, monkeypatch):
    """The final MLP's dense stages run on fp16 two-term splits (k_dense_tc<N, true>) with the 3xTF32 kernel standing by;
    TGNN_DENSE=tf32 keeps them on 3xTF32.  A BatchNorm gain of 1e6 in the first final-MLP layer pushes the inputs of the
    second stage outside the fp16 range: the producers must raise the range flag and the stand-by must redo that stage."""
    if variant == "tf32":
        monkeypatch.setenv("TGNN_DENSE", "tf32")
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(5000, 8, 8, seed=6)
    p = dict(orc.make_params(3, 19, 4, seed=6))
    gold = orc.forward(p, x, ai, af, ci, depth=4, dtype=torch.float64)[:, 0].numpy()
    err = np.abs(run(make_net(p, 3, 19, 4, dev), x, ai, af, ci, dev) - gold).max()
    print(f"final MLP on {variant}: max err {err:.2e}")
    assert err <= TOL
    p["final_mlp.0.mlp.0.batch_norm.weight"] = p["final_mlp.0.mlp.0.batch_norm.weight"] * 1e6
    gold = orc.forward(p, x, ai, af, ci, depth=4, dtype=torch.float64)[:, 0].numpy()
    net = make_net(p, 3, 19, 4, dev)
    err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
    net.check_errors()
    print(f"final MLP on {variant}, BatchNorm gain 1e6: max err {err:.2e}")
    assert err <= TOL


def test_conv_x_transposed_roles_vs_oracle(dev, monkeypatch):
    """k_conv_x (weights as the A operand of mma.sync, gathered rows as B, pairwise channel scatter) only runs in the
    persistent large-graph geometry (> 2 x #SM tiles of 64 rows): a 24k-node lattice with duplicate edges and
    destinations without in-edges, against the fp64 oracle, train and eval."""
    monkeypatch.setenv("TGNN_CONV", "x")
    from tilingnn_b200 import synthetic as syn
    n = 24000
    x, ai, af, ci = syn.lattice_graph(n, 16, 16, seed=8)
    keep = (ai[1] % 41 != 5)
    ai, af = ai[:, keep], af[keep]
    dup = torch.randint(0, ai.shape[1], (5000,), generator=torch.Generator().manual_seed(3))
    ai, af = torch.cat([ai, ai[:, dup]], 1), torch.cat([af, af[dup]], 0)
    p = orc.make_params(3, 19, 3, seed=8)
    for mode in ("train", "eval"):
        q = orc.calibrate_running_stats(p, x, ai, af, ci, depth=3) if mode == "eval" else p
        gold = orc.forward(q, x, ai, af, ci, depth=3, bn_mode=mode, dtype=torch.float64)[:, 0].numpy()
        net = make_net(q, 3, 19, 3, dev, mode)
        err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
        print(f"k_conv_x 24k lattice {mode}: max err {err:.2e}")
        assert net.info()["conv_kernel"] == 5 and err <= TOL


# ---------------------------------------------------------------------------------------------- #
# small-graph latency layout (round 2): cluster-split k_conv_h, k_gin_s, BatchNorm finished by the consumers, parallel branches
@pytest.mark.parametrize("n,deg", [(600, 12), (1900, 8), (4000, 16), (8000, 8)])
def test_small_graph_layouts_vs_oracle(dev, n, deg):
    """n = 600 / 1900 / 4000 / 8000 -> 10 / 30 / 63 / 125 tiles of 64 rows -> thread-block clusters of 8 / 4 / 2 / 1 CTAs per
    tile (148 SMs); k_gin_s below 4096 nodes, k_gin above; every BatchNorm finished in its consumer's prologue.  Ragged graph:
    some destinations lose all their in-edges, duplicates are added.  Against the fp64 oracle, train and eval."""
    from tilingnn_b200 import synthetic as syn
    x, ai, af, ci = syn.lattice_graph(n, deg, deg, seed=5)
    keep = (ai[1] % 37 != 3)
    ai, af = ai[:, keep], af[keep]
    dup = torch.randint(0, ai.shape[1], (n // 4,), generator=torch.Generator().manual_seed(1))
    ai, af = torch.cat([ai, ai[:, dup]], 1), torch.cat([af, af[dup]], 0)
    p = orc.make_params(3, 19, 4, seed=5)
    for mode in ("train", "eval"):
        q = orc.calibrate_running_stats(p, x, ai, af, ci, depth=4) if mode == "eval" else p
        gold = orc.forward(q, x, ai, af, ci, depth=4, bn_mode=mode, dtype=torch.float64)[:, 0].numpy()
        net = make_net(q, 3, 19, 4, dev, mode)
        err = np.abs(run(net, x, ai, af, ci, dev) - gold).max()
        print(f"small-graph layout n={n} {mode}: max err {err:.2e}, launches {net.info()['launches_per_forward']}")
        assert err <= TOL


@pytest.mark.parametrize("var,val,exact", [("TGNN_BRANCHES", "0", True), ("TGNN_BNFIN", "launch", False), ("TGNN_CONV_CLUSTER", "0", False),
                                           ("TGNN_GIN_S", "0", False), ("TGNN_CONV_W16", "0", False)])
def test_small_graph_switches_agree(dev, var, val, exact, monkeypatch):
    """Each small-graph mechanism can be switched off for A/B runs; the scores must not depend on it: bit-identical for the
    side stream (same kernels, same order of arithmetic), <= 2e-5 (eval-BN) where the order of a floating-point sum changes
    (BatchNorm partial rows, cluster / warp partial tiles, gather order of k_gin_s)."""
    z, x, ai, af, ci = load_graph("c1_heart.npz")
    p = load_ckpt("ckpt_30-60-90.npz")
    base = run(make_net(p, x.shape[1], af.shape[1], 20, dev, "eval"), x, ai, af, ci, dev)
    base_t = run(make_net(p, x.shape[1], af.shape[1], 20, dev, "train"), x, ai, af, ci, dev)
    monkeypatch.setenv(var, val)
    alt = run(make_net(p, x.shape[1], af.shape[1], 20, dev, "eval"), x, ai, af, ci, dev)
    alt_t = run(make_net(p, x.shape[1], af.shape[1], 20, dev, "train"), x, ai, af, ci, dev)
    d = np.abs(alt - base).max()
    print(f"{var}={val}: eval-BN max diff {d:.2e}, train-BN max diff {np.abs(alt_t - base_t).max():.2e}")
    if exact:
        assert d == 0.0 and np.array_equal(alt_t, base_t)
    else:
        assert d <= 2e-5       # (eval-BN's collapsed running variances amplify fp32 rounding by ~300; train-BN on this checkpoint amplifies rounding by ~1e3, SURVEY 8c: reported, gated by the tier-3 test)


def test_weight_tables_survive_set_graph_with_the_same_feature_rows(dev):
    """tgnn_set_graph keeps the per-type weight tables when the new graph has the same distinct edge-feature rows (successive
    layouts of one tile set) and rebuilds them when it does not: a handle that walks through A (51 rows), B (same rows, other
    size), C (continuous features: other rows), B again must score each exactly like a fresh handle does."""
    from tilingnn_b200 import synthetic as syn
    p = orc.make_params(3, 19, 3, seed=2)
    graphs = [syn.lattice_graph(900, 8, 8, seed=4), syn.lattice_graph(2500, 8, 8, seed=4),
              syn.lattice_graph(700, 8, 8, seed=4, continuous_features=True), syn.lattice_graph(2500, 8, 8, seed=4)]
    net = make_net(p, 3, 19, 3, dev)
    for k, (x, ai, af, ci) in enumerate(graphs):
        got = run(net, x, ai, af, ci, dev)
        fresh = run(make_net(p, 3, 19, 3, dev), x, ai, af, ci, dev)
        assert np.array_equal(got, fresh), f"graph {k}: a reused handle scores differently from a fresh one"
        gold = orc.forward(p, x, ai, af, ci, depth=3, bn_mode="train", dtype=torch.float64)[:, 0].numpy()
        assert np.abs(got - gold).max() <= TOL
