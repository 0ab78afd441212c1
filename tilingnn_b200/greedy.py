"""The callers around the scoring path: the probabilistic greedy assembly, the
sub-layout it re-scores every round, and the two quality measures of the reference.

Mirrors, in vectorised numpy (no shapely, no Python-per-edge loops):

* ``solve_by_probablistic_greedy`` / ``label_collision_neighbor`` / ``create_solution``
  (/root/reference/util/algorithms.py:18-62, 196-222) -- same round structure, same blend
  ``(prev^(r-1) * p)^(1/r)``, same ``argsort(-p)`` visiting order with the ``break`` at the first already
  labelled node, same ``exp(p - 1) > uniform()`` acceptance drawn from ``numpy.random`` (pass ``rng`` to use
  another stream);
* ``BrickLayout.compute_sub_layout`` (/root/reference/tiling/brick_layout.py:248-286) -- the induced sub-graph
  on the unlabelled nodes, nodes re-indexed in ascending order, edge order preserved;
* ``Losses.calculate_unsupervised_loss`` and ``Losses.solution_score``
  (/root/reference/solver/ml_solver/losses.py:48-148).  ``solution_score`` needs the area of the union of the
  candidate tiles (``BrickLayout.get_super_contour_poly`` = shapely ``unary_union``); ``union_area`` computes it
  exactly with a horizontal slab decomposition (every tile is convex), without the reference's 1e-6 buffer.

The network calls inside go through ``ML_Solver.predict`` -> ``tilingnn_b200.TilinGNN`` -> ``libtgnn.so``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace

import numpy as np

# inputs/config.py:49-51 of the reference
COLLISION_WEIGHT = 1 / math.log(1 + 1e-1)
ALIGN_LENGTH_WEIGHT = 0.02
AVG_AREA_WEIGHT = 1
EPS = 1e-7          # losses.py:10


def _edges(index):
    """[2, E] int64 view of an edge index that may be the reference's empty ``np.array([])``."""
    a = np.asarray(index)
    return a.reshape(2, -1).astype(np.int64, copy=False) if a.size else np.zeros((2, 0), dtype=np.int64)


def _rows(feat, n_edges, like=None):
    a = np.asarray(feat)
    if a.size:
        return a.reshape(n_edges, -1)
    width = np.asarray(like).shape[-1] if like is not None and np.asarray(like).ndim == 2 else 0
    return np.zeros((0, width), dtype=np.float64)


# ------------------------------------------------------------------------------------------------ #
# f2: sub-layout                                                                                    #
# ------------------------------------------------------------------------------------------------ #
def compute_sub_layout(layout, keep, collide_features=True):
    """Induced sub-graph on the nodes ``keep`` (ascending original indices).  Returns ``(sub_layout, keep)``;
    ``keep[i]`` is the reference's ``node_inverse_index[i]`` (brick_layout.py:248-286).
    ``collide_features=False`` leaves the collision edge FEATURES out (an ``[E_c', 0]`` array): the network never
    reads them (TilinGNN.py:51,63) and gathering them is the dominant host cost of a greedy round."""
    keep = np.asarray(keep, dtype=np.int64)
    n = layout.node_feature.shape[0]
    new = -np.ones(n, dtype=np.int64)
    new[keep] = np.arange(len(keep))

    def sub(index, feat, with_feat=True):
        e = _edges(index)
        keep_e = np.flatnonzero((new[e[0]] >= 0) & (new[e[1]] >= 0))
        if not with_feat:
            return new[e[:, keep_e]], np.zeros((len(keep_e), 0))
        return new[e[:, keep_e]], np.take(_rows(feat, e.shape[1], feat), keep_e, axis=0)
    ci, cf = sub(layout.collide_edge_index, layout.collide_edge_features, collide_features)
    ai, af = sub(layout.align_edge_index, layout.align_edge_features)
    out = replace(layout, node_feature=layout.node_feature[keep], collide_edge_index=ci, collide_edge_features=cf,
                  align_edge_index=ai, align_edge_features=af, tiles=np.asarray(layout.tiles)[keep])
    return out, keep


# ------------------------------------------------------------------------------------------------ #
# f1: greedy assembly                                                                               #
# ------------------------------------------------------------------------------------------------ #
@dataclass
class GreedyResult:
    selection: np.ndarray                 # [N] 0/1 float64          (create_solution: temp_sol)
    score: float                          # Losses.solution_score, nan when it cannot be evaluated
    order: list                           # original indices of the selected nodes in selection order
    rounds: int = 0                       # network calls made
    labels: np.ndarray = field(default=None, repr=False)   # [N] int8: 1 selected, 0 knocked out by a collision


def solve_by_probablistic_greedy(ml_solver, origin_layout, rng=None, complete_graph=None, max_rounds=None, trace=None,
                                 sub_layout="mask"):
    """algorithms.py:18-62.  ``ml_solver.predict(layout) -> np.ndarray[N]`` is the only network access.
    ``trace`` (optional list) receives one ``(round, node, accept threshold exp(p-1), uniform draw, accepted)`` tuple per
    visited node -- the tests use it to locate the first decision at which two score sources part ways.
    ``sub_layout``: "mask" scores a round's sub-layout through ``ml_solver.predict_sub_layout(origin_layout, keep)`` when the
    solver has it (the origin layout stays resident on the GPU, a node mask selects the sub-graph); "reindex" always builds
    the re-indexed sub-layout as the reference does (brick_layout.py:248-286) and calls ``predict`` on it."""
    masked_predict = None
    if sub_layout == "mask" and getattr(ml_solver, "supports_node_mask", False):
        masked_predict = ml_solver.predict_sub_layout
    rng = np.random if rng is None else rng
    n = origin_layout.node_feature.shape[0]
    col = _edges(origin_layout.collide_edge_index)
    order_by_src = np.argsort(col[0], kind="stable")                      # neighbours of i = col[1][col[0] == i]
    nbr = col[1][order_by_src]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(col[0], minlength=n), out=ptr[1:])

    label = -np.ones(n, dtype=np.int8)                                    # -1 = unlabelled_nodes
    saved = np.ones(n, dtype=np.float64)                                  # unlabelled_nodes[key]
    order = []
    round_cnt = 1
    while (label < 0).any():
        if max_rounds is not None and round_cnt > max_rounds:
            raise RuntimeError(f"greedy assembly did not finish in {max_rounds} rounds")
        keep = np.flatnonzero(label < 0)
        if masked_predict is not None:
            node_re_index = keep
            prob = np.asarray(masked_predict(origin_layout, keep))
        else:
            temp_layout, node_re_index = compute_sub_layout(origin_layout, keep, collide_features=False)
            prob = np.asarray(ml_solver.predict(temp_layout))
        previous_prob = saved[keep]
        prob_per_node = np.power(np.power(previous_prob, round_cnt - 1) * prob, 1 / round_cnt)
        saved[keep] = prob_per_node
        # the visiting loop runs on plain Python scalars (a numpy scalar operation costs ~10x a float one); the
        # acceptance threshold exp(p - 1) is evaluated by numpy for the whole round at once, elementwise as the reference
        accept_at = np.exp((prob_per_node - 1) * 1.0).tolist()
        origin_of = node_re_index.tolist()
        uniform = rng.uniform
        for idx in np.argsort(-prob_per_node).tolist():
            origin_idx = origin_of[idx]
            if label[origin_idx] >= 0:                                    # collision handling: stop this round
                break
            u = uniform()
            if trace is not None:
                trace.append((round_cnt, origin_idx, accept_at[idx], u, accept_at[idx] > u))
            if accept_at[idx] > u:
                label[origin_idx] = 1
                order.append(origin_idx)
                adj = nbr[ptr[origin_idx]:ptr[origin_idx + 1]]            # label_collision_neighbor (:196-207)
                label[adj[label[adj] < 0]] = 0
        round_cnt += 1
    selection = (label == 1).astype(np.float64)
    g = complete_graph if complete_graph is not None else getattr(ml_solver, "complete_graph", None)
    score = solution_score(selection, origin_layout, g) if g is not None else float("nan")
    return GreedyResult(selection, score, order, round_cnt - 1, label)


# ------------------------------------------------------------------------------------------------ #
# f4: quality measures                                                                              #
# ------------------------------------------------------------------------------------------------ #
def calculate_unsupervised_loss(probs, node_feature, collide_edge_index, adj_edges_index, adj_edge_features):
    """losses.py:48-116 for ``probs`` [N, M].  Returns ``(min loss, argmin, losses[M])``."""
    probs = np.asarray(probs, dtype=np.float64).reshape(len(probs), -1)
    x = np.asarray(node_feature, dtype=np.float64)
    col, adj = _edges(collide_edge_index), _edges(adj_edges_index)
    losses = []
    for sol in range(probs.shape[1]):
        p = probs[:, sol]
        loss_ave_area = math.log(max(float(np.mean(x[:, -1] * p)), EPS))
        loss_feasibility = 0.0
        if col.shape[1] > 0:
            pp = np.clip(p[col[0]] * p[col[1]], EPS, 1 - EPS)
            loss_feasibility = float(np.log(1 - pp).sum() / col.shape[1])
        loss_align_length = 0.0
        if adj.shape[1] > 0:
            lengths = np.asarray(adj_edge_features, dtype=np.float64)[:, 1]
            pp = np.maximum(p[adj[0]] * p[adj[1]] * lengths, EPS)
            loss_align_length = float((np.log(pp) / math.log(10)).sum() / adj.shape[1])
        assert loss_feasibility <= 0 and loss_ave_area <= 0 and loss_align_length <= 0
        loss = ((1 - AVG_AREA_WEIGHT * loss_ave_area) * (1 - COLLISION_WEIGHT * loss_feasibility)
                * (1 - ALIGN_LENGTH_WEIGHT * loss_align_length))
        assert loss >= 1.0
        losses.append(loss)
    losses = np.asarray(losses)
    return float(losses.min()), int(np.argmin(losses)), losses


def ring_perimeter(ring):
    r = np.asarray(ring, dtype=np.float64)
    d = r - np.roll(r, 1, axis=0)               # closed rings contribute a zero-length segment for the repeated vertex
    return float(np.sqrt((d * d).sum(axis=1)).sum())


def union_area(rings, tol=1e-9):
    """Exact area of the union of CONVEX polygons (closed or open rings, any orientation).

    Horizontal slabs between consecutive critical ordinates (vertices and proper edge crossings): inside a slab
    no two edges cross, so the covered length of a horizontal line is linear in y and the trapezoid rule on the
    two slab ends is exact."""
    polys = []
    for r in rings:
        r = np.asarray(r, dtype=np.float64)
        if len(r) > 1 and np.allclose(r[0], r[-1]):
            r = r[:-1]
        if len(r) >= 3:
            polys.append(r)
    if not polys:
        return 0.0
    pid = np.concatenate([np.full(len(p), i) for i, p in enumerate(polys)])
    a = np.concatenate(polys)
    b = np.concatenate([np.roll(p, -1, axis=0) for p in polys])
    keep = np.abs(a[:, 1] - b[:, 1]) > tol                                # horizontal edges bound no slab interval
    a, b, pid = a[keep], b[keep], pid[keep]
    swap = a[:, 1] > b[:, 1]
    a[swap], b[swap] = b[swap].copy(), a[swap].copy()                     # a = lower end
    ys = [a[:, 1], b[:, 1]]
    d = b - a
    m = len(a)
    blk = max(1, (1 << 22) // max(m, 1))
    for s in range(0, m, blk):                                            # proper crossings, block-wise
        e = slice(s, min(m, s + blk))
        da, aa = d[e, None, :], a[e, None, :]
        den = da[..., 0] * d[None, :, 1] - da[..., 1] * d[None, :, 0]
        w = a[None, :, :] - aa
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (w[..., 0] * d[None, :, 1] - w[..., 1] * d[None, :, 0]) / den
            u = (w[..., 0] * da[..., 1] - w[..., 1] * da[..., 0]) / den
        hit = (np.abs(den) > 1e-14) & (t > tol) & (t < 1 - tol) & (u > tol) & (u < 1 - tol)
        ys.append((aa[..., 1] + t * da[..., 1])[hit])
    y = np.unique(np.round(np.concatenate(ys) / tol) * tol)
    y = y[np.concatenate([[True], np.diff(y) > tol])]
    total = 0.0
    inv_dy = d[:, 0] / d[:, 1]
    for y0, y1 in zip(y[:-1], y[1:]):
        span = (a[:, 1] <= y0 + tol) & (b[:, 1] >= y1 - tol)              # edges crossing the whole slab
        if not span.any():
            continue
        p = pid[span]
        x0 = a[span, 0] + (y0 - a[span, 1]) * inv_dy[span]
        x1 = a[span, 0] + (y1 - a[span, 1]) * inv_dy[span]
        srt = np.argsort(p, kind="stable")
        p, x0, x1 = p[srt], x0[srt], x1[srt]
        first = np.flatnonzero(np.concatenate([[True], p[1:] != p[:-1]]))
        lo0, hi0 = np.minimum.reduceat(x0, first), np.maximum.reduceat(x0, first)
        lo1, hi1 = np.minimum.reduceat(x1, first), np.maximum.reduceat(x1, first)
        total += 0.5 * (y1 - y0) * (_covered(lo0, hi0) + _covered(lo1, hi1))
    return float(total)


def _covered(lo, hi):
    """Length of the union of the intervals [lo_i, hi_i]."""
    o = np.argsort(lo, kind="stable")
    lo, hi = lo[o], hi[o]
    reach = np.maximum.accumulate(hi)
    start = np.concatenate([[True], lo[1:] > reach[:-1]])
    ends = np.concatenate([reach[:-1][start[1:]], reach[-1:]])
    return float((ends - lo[start]).sum())


def solution_score(predict, brick_layout, complete_graph):
    """losses.py:119-148: ``AVG_AREA_WEIGHT * filled area ratio + ALIGN_LENGTH_WEIGHT * aligned length ratio``."""
    g = complete_graph
    predict = np.asarray(predict, dtype=np.float64)
    x = np.asarray(brick_layout.node_feature, dtype=np.float64)
    tiles = np.asarray(brick_layout.tiles)
    contour_area = getattr(brick_layout, "_super_contour_area", None)
    if contour_area is None:
        contour_area = union_area([g.tile_rings[t] for t in tiles])
        try:
            object.__setattr__(brick_layout, "_super_contour_area", contour_area)
        except Exception:
            pass
    filled_area = float(predict.dot(x[:, -1] * g.max_area) / contour_area)
    assert -1e-7 <= filled_area <= 1 + 1e-7
    adj = _edges(brick_layout.align_edge_index)
    loss_align_length = 0.0
    if adj.shape[1] > 0:
        lengths = np.asarray(brick_layout.align_edge_features, dtype=np.float64)[:, 1] * g.max_align_length
        loss_align_length = float((predict[adj[0]] * predict[adj[1]]).dot(lengths))
    perim = getattr(brick_layout, "_tile_perimeters", None)
    if perim is None:
        perim = np.asarray([ring_perimeter(g.tile_rings[t]) for t in tiles])
        try:
            object.__setattr__(brick_layout, "_tile_perimeters", perim)
        except Exception:
            pass
    all_edge_length = float(perim[predict == 1].sum())
    if all_edge_length == 0:
        return float("nan")
    ratio = loss_align_length / all_edge_length
    assert -1e-7 < ratio < 1 + 1e-7
    return float(AVG_AREA_WEIGHT * filled_area + ALIGN_LENGTH_WEIGHT * ratio)
