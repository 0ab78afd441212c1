"""The callers around the scoring path against golden vectors produced by the reference's
own source text (tests/golden/make_greedy_golden.py): greedy assembly decisions, sub-layout re-indexing, the
unsupervised loss; plus the shapely-free union area / solution score on cases with known answers."""
import os

import numpy as np
import pytest

from _util import GOLDEN, fake_predict
from tilingnn_b200 import greedy
from tilingnn_b200.tile_graph_io import SuperGraph


@pytest.fixture(scope="module")
def heart():
    z = dict(np.load(os.path.join(GOLDEN, "greedy_heart.npz")))
    af = z["align_feat_rows"][z["align_feat_id"].astype(np.int64)]
    ci = z["collide_edge_index"].astype(np.int64)
    sg = SuperGraph(z["node_feature"], ci, np.zeros((ci.shape[1], af.shape[1])), z["align_edge_index"].astype(np.int64),
                    af, z["tiles"])
    return z, sg


class FakeSolver:
    complete_graph = None

    def __init__(self):
        self.sizes = []

    def predict(self, layout):
        self.sizes.append(layout.node_feature.shape[0])
        if np.size(layout.collide_edge_index) == 0 or np.size(layout.align_edge_index) == 0:
            return np.ones(layout.node_feature.shape[0], dtype=np.float32)
        return fake_predict(layout.node_feature, layout.collide_edge_index, layout.align_edge_index)


def test_greedy_makes_the_reference_decisions(heart):
    z, sg = heart
    solver = FakeSolver()
    res = greedy.solve_by_probablistic_greedy(solver, sg, rng=np.random.RandomState(2))
    assert solver.sizes == z["round_sizes"].tolist()                    # same sub-layout every round
    assert res.order == z["order"].tolist()                             # same tiles in the same order
    assert np.array_equal(res.selection, z["selection"])
    lab = np.full(len(res.labels), -1)
    lab[z["labelled_keys"]] = z["labelled_vals"]
    assert np.array_equal(res.labels, lab)
    # and with numpy's global stream, which is what the reference draws from
    np.random.seed(2)
    assert greedy.solve_by_probablistic_greedy(FakeSolver(), sg).order == z["order"].tolist()


def test_greedy_through_the_node_mask_path_makes_the_reference_decisions(heart):
    """a solver that offers ``predict_sub_layout(origin, keep)`` (the GPU solver does: node mask on the resident graph) is
    handed the kept node ids instead of a re-indexed layout; the loop around it must make the same decisions."""
    z, sg = heart

    class MaskSolver(FakeSolver):
        supports_node_mask = True

        def predict_sub_layout(self, origin, keep):
            sub, back = greedy.compute_sub_layout(origin, keep, collide_features=False)
            assert np.array_equal(back, keep)
            return self.predict(sub)
    solver = MaskSolver()
    res = greedy.solve_by_probablistic_greedy(solver, sg, rng=np.random.RandomState(2), sub_layout="mask")
    assert solver.sizes == z["round_sizes"].tolist() and res.order == z["order"].tolist()
    assert np.array_equal(res.selection, z["selection"])
    ref = greedy.solve_by_probablistic_greedy(MaskSolver(), sg, rng=np.random.RandomState(2), sub_layout="reindex")
    assert ref.order == res.order


def test_greedy_solution_is_a_maximal_independent_set(heart):
    _, sg = heart
    res = greedy.solve_by_probablistic_greedy(FakeSolver(), sg, rng=np.random.RandomState(7))
    sel = res.selection.astype(bool)
    ci = sg.collide_edge_index
    assert not (sel[ci[0]] & sel[ci[1]]).any(), "two selected tiles collide"
    blocked = np.zeros(len(sel), bool)
    blocked[ci[1][sel[ci[0]]]] = True
    assert (sel | blocked).all(), "an unselected tile collides with no selected tile"
    assert (res.labels >= 0).all() and res.rounds >= 1


def test_sub_layout_matches_the_reference(heart):
    z, sg = heart
    keep = z["sub_keep"]
    sub, inv = greedy.compute_sub_layout(sg, keep)
    assert np.array_equal(inv, keep)
    assert np.array_equal(sub.node_feature, z["sub_node_feature"])
    assert np.array_equal(sub.collide_edge_index, z["sub_collide"])
    assert np.array_equal(sub.align_edge_index, z["sub_align"])
    assert np.array_equal(sub.align_edge_features[:, 1], z["sub_align_feat_col1"])
    assert np.array_equal(sub.tiles, sg.tiles[keep])
    empty, _ = greedy.compute_sub_layout(sg, np.array([0]))            # one node: no edges left
    assert empty.collide_edge_index.shape == (2, 0) and empty.align_edge_features.shape[0] == 0


def test_unsupervised_loss_matches_the_reference(heart):
    z, sg = heart
    loss, arg, losses = greedy.calculate_unsupervised_loss(z["loss_probs"], sg.node_feature, sg.collide_edge_index,
                                                           sg.align_edge_index, sg.align_edge_features)
    assert np.allclose(losses, z["losses"], rtol=1e-12) and arg == int(z["loss_min_index"]) and loss == losses.min()
    # no collision edges / no adjacency edges (losses.py:69-70, 82-83)
    e = np.array([])
    l0 = greedy.calculate_unsupervised_loss(z["loss_probs"][:, :1], sg.node_feature, e, e, e)[0]
    assert l0 == pytest.approx(1 - np.log(np.mean(sg.node_feature[:, -1] * z["loss_probs"][:, 0])))


def test_union_area_known_answers():
    sq = lambda x, y, s: np.array([[x, y], [x + s, y], [x + s, y + s], [x, y + s]], dtype=float)
    assert greedy.union_area([sq(0, 0, 1)]) == pytest.approx(1.0)
    assert greedy.union_area([sq(0, 0, 1), sq(0.5, 0.5, 1)]) == pytest.approx(1.75)
    assert greedy.union_area([sq(0, 0, 1), sq(1, 0, 1), sq(0, 0, 1)[::-1]]) == pytest.approx(2.0)      # shared edge, duplicate, cw
    tri = np.array([[0, 0], [2, 0], [1, 2], [0, 0]], dtype=float)                                          # closed ring
    assert greedy.union_area([tri, sq(0.5, 0, 1)]) == pytest.approx(2.0)            # square inside the triangle
    assert greedy.union_area([tri, sq(-0.5, 0, 1)]) == pytest.approx(2.75)         # overlap = int_0^1 (0.5 - y/2) dy
    rng = np.random.default_rng(0)
    polys = []
    for _ in range(25):
        c, r, k = rng.uniform(0, 4, 2), rng.uniform(0.3, 1.0), rng.integers(3, 7)
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        polys.append(c + r * np.stack([np.cos(ang), np.sin(ang)], 1))
    gx, gy = np.meshgrid(np.linspace(-1.2, 5.2, 1601), np.linspace(-1.2, 5.2, 1601))
    pts = np.stack([gx.ravel(), gy.ravel()], 1)
    inside = np.zeros(len(pts), bool)
    for p in polys:
        a, e = p, np.roll(p, -1, 0) - p
        cr = e[None, :, 0] * (pts[:, None, 1] - a[None, :, 1]) - e[None, :, 1] * (pts[:, None, 0] - a[None, :, 0])
        inside |= (cr >= 0).all(1) | (cr <= 0).all(1)
    assert greedy.union_area(polys) == pytest.approx(inside.mean() * 6.4 ** 2, rel=5e-3)


def test_solution_score_on_a_toy_tiling():
    """two unit squares side by side + one overlapping both; select the two disjoint ones."""
    from types import SimpleNamespace
    sq = lambda x: np.array([[x, 0], [x + 1, 0], [x + 1, 1], [x, 1], [x, 0]], dtype=float)
    g = SimpleNamespace(tile_rings=[sq(0), sq(1), sq(0.5)], max_area=1.0, max_align_length=1.0)
    nf = np.array([[1, 1.0]] * 3)
    ai = np.array([[0, 1], [1, 0]])
    af = np.array([[0, 1.0, 1], [0, 1.0, 1]])
    ci = np.array([[0, 2, 1, 2], [2, 0, 2, 1]])
    lay = SuperGraph(nf, ci, np.zeros((4, 3)), ai, af, np.arange(3))
    s = greedy.solution_score(np.array([1.0, 1.0, 0.0]), lay, g)
    # filled 2 / union 2 = 1; aligned length: both directed edges count (1 + 1) / perimeters (4 + 4)
    assert s == pytest.approx(1 * 1.0 + 0.02 * (2 / 8))
    assert np.isnan(greedy.solution_score(np.zeros(3), lay, g))
