"""Shapely-free readers for the reference's on-disk formats (host side, numpy only).

These are the data formats either side of the scoring path:

* complete-graph pickles ``data/<env>/complete_graph_ring*.pkl`` written by
  ``TileGraph.save_current_state`` (/root/reference/tiling/tile_graph.py:296-308)
  -- the pickled ``shapely`` polygons are decoded from their WKB payload, so
  neither ``shapely`` nor the reference package has to be importable;
* silhouette contour files ``silhouette/*.txt`` read by ``load_polygons``
  (/root/reference/util/shape_processor.py:18-32);
* the silhouette -> super-graph crop of
  ``crop_multiple_layouts_from_contour`` / ``shape_transform`` /
  ``get_all_placement_in_polygon`` / ``generate_brick_layout_data``
  (/root/reference/tiling/tile_factory.py:162-201, 222-243, 37-47 and
  /root/reference/util/data_util.py:164-204).  ``contain(poly, tile)`` is
  ``|area(poly ∩ tile) - area(tile)| < 1e-6`` (/root/reference/util/algo_util.py:143-144);
  every tile is convex, so clipping each contour ring by the tile
  (Sutherland-Hodgman) and the shoelace formula give that area exactly.
"""
from __future__ import annotations

import itertools
import pickle
import struct
from dataclasses import dataclass, field

import numpy as np


# --------------------------------------------------------------------------- #
# pickle reader                                                               #
# --------------------------------------------------------------------------- #

class _Opaque:
    """Stand-in for pickled ``tiling.tile.Tile`` / ``shapely`` objects."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.state = state


# The graph pickles are external data files.  Only the few globals they actually need are resolved; anything else
# (os.system, builtins.eval, ...) raises instead of being imported -- a crafted file cannot run code through this loader.
_SAFE_GLOBALS = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("builtins", "list"), ("builtins", "dict"), ("builtins", "tuple"), ("builtins", "set"), ("builtins", "frozenset"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "complex"), ("builtins", "bool"), ("builtins", "str"),
    ("builtins", "bytes"), ("builtins", "bytearray"), ("builtins", "slice"), ("builtins", "range"), ("builtins", "object"),
    ("copyreg", "_reconstructor"),
}


class _GraphUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("shapely", "tiling"):
            return type(name, (_Opaque,), {})
        if (module, name) in _SAFE_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"graph pickle references {module}.{name}, which is not on the allow-list")


def _wkb_polygon_exterior(buf: bytes) -> np.ndarray:
    """Exterior ring (closed, [n,2] float64) of a WKB polygon."""
    bo = "<" if buf[0] == 1 else ">"
    gtype, nrings = struct.unpack_from(bo + "II", buf, 1)
    if gtype & 0xFF != 3 or nrings < 1:
        raise ValueError(f"not a WKB polygon (type {gtype}, rings {nrings})")
    (npts,) = struct.unpack_from(bo + "I", buf, 9)
    dims = 3 if (gtype & 0x80000000 or gtype // 1000 == 1) else 2
    pts = np.frombuffer(buf, dtype=np.dtype(bo + "f8"), count=npts * dims, offset=13)
    return pts.reshape(npts, dims)[:, :2].copy()


def shoelace_area(ring: np.ndarray) -> float:
    x, y = ring[:, 0], ring[:, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))))


@dataclass
class CompleteGraph:
    """The fields of the reference ``TileGraph`` that the scoring path needs."""
    tile_rings: list                 # per tile: closed exterior ring [n,2]
    tile_ids: np.ndarray             # [N] int, prototype-tile id (Tile.id)
    tile_areas: np.ndarray           # [N] float64
    adj_edges: np.ndarray            # [E_a,2] int64 (u,v); both directions present
    colli_edges: np.ndarray          # [E_c,2] int64
    adj_features: np.ndarray         # [E_a,D_e] float64, column 1 NOT yet / max_align_length
    colli_features: np.ndarray       # [E_c,D_e] float64
    max_area: float
    max_align_length: float
    align_start_index: int
    total_feature_dim: int
    tile_type_count: int
    _edge_pos: dict = field(default_factory=dict, repr=False)

    @property
    def num_nodes(self):
        return len(self.tile_ids)


def load_complete_graph(path, tile_type_count=None) -> CompleteGraph:
    """Read ``complete_graph_ring*.pkl`` (tile_graph.py:310-332) without shapely."""
    with open(path, "rb") as f:
        d = _GraphUnpickler(f).load()
    rings, ids = [], []
    for t in d["tiles"]:
        st = t.state
        poly = st["tile_poly"]
        rings.append(_wkb_polygon_exterior(poly.state))
        ids.append(int(st["id"]))
    ids = np.asarray(ids, dtype=np.int64)
    areas = np.asarray([shoelace_area(r[:-1]) for r in rings])
    ef = d["edges_features"]
    adj = np.asarray(d["adj_edges"], dtype=np.int64).reshape(-1, 2)
    col = np.asarray(d["colli_edges"], dtype=np.int64).reshape(-1, 2)
    adj_f = np.asarray([ef[u][v] for u, v in adj], dtype=np.float64)
    col_f = np.asarray([ef[u][v] for u, v in col], dtype=np.float64)
    total_dim = int(d["align_start_index"]) + len(d["unique_adj_features"])
    if tile_type_count is None:
        tile_type_count = int(ids.max()) + 1
    return CompleteGraph(rings, ids, areas, adj, col, adj_f, col_f,
                         float(d["max_area"]), float(d["max_align_length"]),
                         int(d["align_start_index"]), total_dim, int(tile_type_count))


# --------------------------------------------------------------------------- #
# silhouettes and cropping                                                    #
# --------------------------------------------------------------------------- #

def load_polygons(filename):
    """``load_polygons`` (shape_processor.py:18-32): line 0 exterior, others holes."""
    def parse(line):
        return np.asarray([[float(w.split(" ")[0]), float(w.split(" ")[1])]
                           for w in line.strip().split(",")])
    lines = [ln for ln in open(filename) if ln.strip()]
    return parse(lines[0]), [parse(ln) for ln in lines[1:]]


def _signed_area_centroid(ring):
    x, y = ring[:, 0], ring[:, 1]
    xn, yn = np.roll(x, -1), np.roll(y, -1)
    cr = x * yn - xn * y
    a = 0.5 * cr.sum()
    cx = ((x + xn) * cr).sum() / (6 * a)
    cy = ((y + yn) * cr).sum() / (6 * a)
    return a, np.array([cx, cy])


def _open_ring(r):
    r = np.asarray(r, dtype=np.float64)
    if len(r) > 1 and np.allclose(r[0], r[-1]):
        r = r[:-1]
    return r


def polygon_centroid(exterior, interiors):
    a, c = _signed_area_centroid(_open_ring(exterior))
    tot_a, tot_c = abs(a), c * abs(a)
    for h in interiors:
        ah, ch = _signed_area_centroid(_open_ring(h))
        tot_a -= abs(ah)
        tot_c -= ch * abs(ah)
    return tot_c / tot_a


def _clip_by_convex(subject, clip):
    """Sutherland-Hodgman: clip open ring ``subject`` by CONVEX open ring ``clip``."""
    a_clip, _ = _signed_area_centroid(clip)
    if a_clip < 0:
        clip = clip[::-1]
    out = subject
    m = len(clip)
    for i in range(m):
        if len(out) == 0:
            break
        p, q = clip[i], clip[(i + 1) % m]
        ex, ey = q[0] - p[0], q[1] - p[1]
        d = ex * (out[:, 1] - p[1]) - ey * (out[:, 0] - p[0])     # >= 0: inside (left of p->q)
        nxt = np.roll(out, -1, axis=0)
        dn = np.roll(d, -1)
        pieces = []
        for k in range(len(out)):
            s_in, e_in = d[k] >= 0, dn[k] >= 0
            if s_in:
                pieces.append(out[k])
            if s_in != e_in:
                t = d[k] / (d[k] - dn[k])
                pieces.append(out[k] + t * (nxt[k] - out[k]))
        out = np.asarray(pieces).reshape(-1, 2)
    return out


def intersection_area_with_convex(exterior, interiors, tile_ring):
    tile = _open_ring(tile_ring)
    c = _clip_by_convex(_open_ring(exterior), tile)
    a = shoelace_area(c) if len(c) >= 3 else 0.0
    for h in interiors:
        ch = _clip_by_convex(_open_ring(h), tile)
        if len(ch) >= 3:
            a -= shoelace_area(ch)
    return a


def _clip_area_batch(subject, tiles):
    """Area of ``subject`` (open ring [n,2], any orientation) clipped by each CONVEX tile of ``tiles`` [B,m,2] (open
    rings): Sutherland-Hodgman for all tiles at once on padded arrays -- the same arithmetic as ``_clip_by_convex`` +
    ``shoelace_area`` per tile, without the Python loop over tiles and vertices."""
    tiles = np.asarray(tiles, dtype=np.float64)
    B, m = tiles.shape[:2]
    x, y = tiles[:, :, 0], tiles[:, :, 1]
    signed = 0.5 * (x * np.roll(y, -1, axis=1) - np.roll(x, -1, axis=1) * y).sum(axis=1)
    tiles = np.where((signed < 0)[:, None, None], tiles[:, ::-1], tiles)                 # counter-clockwise clip rings
    n = len(subject)
    pts = np.broadcast_to(np.asarray(subject, dtype=np.float64), (B, n, 2)).copy()
    cnt = np.full(B, n, dtype=np.int64)
    rows = np.arange(B)[:, None]
    for i in range(m):
        L = pts.shape[1]
        if L == 0:
            break
        p, q = tiles[:, i], tiles[:, (i + 1) % m]
        ex, ey = (q[:, 0] - p[:, 0])[:, None], (q[:, 1] - p[:, 1])[:, None]
        ar = np.arange(L)[None, :]
        valid = ar < cnt[:, None]
        d = ex * (pts[:, :, 1] - p[:, 1, None]) - ey * (pts[:, :, 0] - p[:, 0, None])     # >= 0: inside
        nxt_i = (ar + 1) % np.maximum(cnt, 1)[:, None]
        nxt = pts[rows, nxt_i]
        dn = d[rows, nxt_i]
        s_in, e_in = d >= 0, dn >= 0
        emit1 = s_in & valid
        emit2 = (s_in != e_in) & valid
        with np.errstate(divide="ignore", invalid="ignore"):
            t = d / (d - dn)
            inter = pts + t[:, :, None] * (nxt - pts)                                     # only read where emit2
        c = emit1.astype(np.int64) + emit2.astype(np.int64)
        pos = np.cumsum(c, axis=1) - c
        new_cnt = c.sum(axis=1)
        out = np.zeros((B, max(int(new_cnt.max()), 1), 2))
        bb = np.broadcast_to(rows, (B, L))
        out[bb[emit1], pos[emit1]] = pts[emit1]
        pos2 = pos + emit1
        out[bb[emit2], pos2[emit2]] = inter[emit2]
        pts, cnt = out, new_cnt
    L = pts.shape[1]
    if L == 0:
        return np.zeros(B)
    ar = np.arange(L)[None, :]
    valid = ar < cnt[:, None]
    nxt_i = (ar + 1) % np.maximum(cnt, 1)[:, None]
    nx = pts[rows, nxt_i]
    cross = np.where(valid, pts[:, :, 0] * nx[:, :, 1] - nx[:, :, 0] * pts[:, :, 1], 0.0)
    area = 0.5 * np.abs(cross.sum(axis=1))
    return np.where(cnt >= 3, area, 0.0)


def intersection_areas_with_convex(exterior, interiors, tile_rings):
    """``intersection_area_with_convex`` for many tiles (rings grouped by vertex count and clipped in batches)."""
    open_rings = [_open_ring(r) for r in tile_rings]
    out = np.zeros(len(open_rings))
    ext = _open_ring(exterior)
    holes = [_open_ring(h) for h in interiors]
    by_m = {}
    for i, r in enumerate(open_rings):
        by_m.setdefault(len(r), []).append(i)
    for m, idx in by_m.items():
        tiles = np.stack([open_rings[i] for i in idx])
        a = _clip_area_batch(ext, tiles)
        for h in holes:
            a = a - _clip_area_batch(h, tiles)
        out[idx] = a
    return out


def graph_bound(g: CompleteGraph):
    allpts = np.concatenate(g.tile_rings, axis=0)
    return allpts[:, 0].min(), allpts[:, 0].max(), allpts[:, 1].min(), allpts[:, 1].max()


def shape_transform(g, exterior, interiors, margin_padding_ratio, rotate_angle, x_delta, y_delta):
    """``shape_transform`` (tile_factory.py:222-243): scale to the graph, centre, rotate, shift."""
    exterior = np.asarray(exterior, dtype=np.float64)
    lo, hi = exterior.min(axis=0), exterior.max(axis=0)
    max_axis = max(hi[0] - lo[0], hi[1] - lo[1])
    x_min, x_max, y_min, y_max = graph_bound(g)
    center = np.array([(x_max + x_min) / 2, (y_max + y_min) / 2])
    base_diameter = min(x_max - x_min, y_max - y_min)
    s = base_diameter * margin_padding_ratio / max_axis
    ext = exterior * s
    ints = [np.asarray(h, dtype=np.float64) * s for h in interiors]
    c = polygon_centroid(ext, ints)
    ext, ints = ext - c, [h - c for h in ints]
    if rotate_angle != 0.0:
        # shapely.affinity.rotate(origin="centroid"): the centroid is now the origin
        th = np.deg2rad(rotate_angle)
        rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        ext, ints = ext @ rot.T, [h @ rot.T for h in ints]
    shift = center + np.array([x_delta, y_delta])
    return base_diameter, ext + shift, [h + shift for h in ints]


@dataclass
class SuperGraph:
    """The five arrays ``BrickLayout`` hands to the network (brick_layout.py:242-246)."""
    node_feature: np.ndarray             # [n, tile_type_count+1] float64
    collide_edge_index: np.ndarray       # [2, e_c] int64 (row0 = u, row1 = v)
    collide_edge_features: np.ndarray    # [e_c, D_e] float64
    align_edge_index: np.ndarray         # [2, e_a] int64
    align_edge_features: np.ndarray      # [e_a, D_e] float64, col 1 / max_align_length
    tiles: np.ndarray                    # [n] indices into the complete graph


def super_graph_from_tiles(g: CompleteGraph, tiles_super_set) -> SuperGraph:
    """``get_all_placement_in_polygon`` edge filter + ``generate_brick_layout_data``."""
    tiles = np.asarray(tiles_super_set, dtype=np.int64)
    inv = -np.ones(g.num_nodes, dtype=np.int64)
    inv[tiles] = np.arange(len(tiles))
    ka = (inv[g.adj_edges[:, 0]] >= 0) & (inv[g.adj_edges[:, 1]] >= 0)
    kc = (inv[g.colli_edges[:, 0]] >= 0) & (inv[g.colli_edges[:, 1]] >= 0)
    adj_f = g.adj_features[ka].copy()
    if len(adj_f) > 0:
        adj_f[:, 1] = adj_f[:, 1] / g.max_align_length                 # data_util.py:167-169
    nf = np.zeros((len(tiles), g.tile_type_count + 1))
    nf[np.arange(len(tiles)), g.tile_ids[tiles]] = 1                    # data_util.py:185-189
    nf[:, -1] = g.tile_areas[tiles] / g.max_area
    return SuperGraph(nf, inv[g.colli_edges[kc]].T.copy(), g.colli_features[kc].copy(),
                      inv[g.adj_edges[ka]].T.copy(), adj_f, tiles)


def complete_super_graph(g: CompleteGraph) -> SuperGraph:
    return super_graph_from_tiles(g, np.arange(g.num_nodes))


def crop_from_contour(g, exterior, interiors, margin_padding_ratio=0.5, rotate_angle=0.0,
                      x_delta=0.0, y_delta=0.0) -> SuperGraph:
    _, ext, ints = shape_transform(g, exterior, interiors, margin_padding_ratio,
                                   rotate_angle, x_delta, y_delta)
    lo, hi = ext.min(axis=0), ext.max(axis=0)
    bb = g._edge_pos.get("tile_bbox")                                   # per-graph cache: [N,4] xmin, ymin, xmax, ymax
    if bb is None:
        bb = np.asarray([[*r.min(axis=0), *r.max(axis=0)] for r in g.tile_rings])
        g._edge_pos["tile_bbox"] = bb
    inside = (bb[:, 0] >= lo[0] - 1e-9) & (bb[:, 1] >= lo[1] - 1e-9) & (bb[:, 2] <= hi[0] + 1e-9) & (bb[:, 3] <= hi[1] + 1e-9)
    cand = np.flatnonzero(inside).tolist()                              # bbox reject: the others cannot be contained
    if not cand:
        return super_graph_from_tiles(g, [])
    areas = intersection_areas_with_convex(ext, ints, [g.tile_rings[i] for i in cand])
    keep = [i for i, a in zip(cand, areas) if abs(a - g.tile_areas[i]) < 1e-6]      # algo_util.py:143-144
    return super_graph_from_tiles(g, keep)


def tile_movement_delta(g, movement_delta_ratio):
    """``get_tile_movement_delta`` (tile_factory.py:204-209)."""
    r = g.tile_rings[0]
    d = min(r[:, 0].max() - r[:, 0].min(), r[:, 1].max() - r[:, 1].min())
    return np.asarray(movement_delta_ratio, dtype=np.float64) * d


def crop_multiple_layouts_from_contour(exterior, interiors, g, start_angle=0.0, end_angle=60.0,
                                       num_of_angle=1, movement_delta_ratio=(0,),
                                       margin_padding_ratios=(0.2,)):
    """``crop_multiple_layouts_from_contour`` (tile_factory.py:162-201), geometry only."""
    out = []
    deltas = tile_movement_delta(g, movement_delta_ratio)
    for margin in margin_padding_ratios:
        for ang in np.linspace(start_angle, end_angle, num_of_angle):
            for dx, dy in itertools.product(deltas, deltas):
                sg = crop_from_contour(g, exterior, interiors, margin, float(ang), dx, dy)
                if sg.node_feature.shape[0] == 0:
                    continue
                out.append(sg)
    return out
