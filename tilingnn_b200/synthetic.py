"""Seeded synthetic super-graphs with the statistics of the reference's tile graphs
(SURVEY.md §8d / Appendix B): nodes in row-major order on a sqrt(N) x sqrt(N) lattice, each node's
in-neighbours taken from a fixed symmetric stencil (adjacency = the nearest ``deg_adj`` offsets,
collision = the next ``deg_col`` offsets; edges that would leave the lattice are dropped, so the
degree is exact in the interior and the average slightly lower), both edge sets symmetric, no self
loops, no duplicates -> banded structure, halo ~ sqrt(N) under node-range sharding.

Adjacency edge features mimic /root/reference/tiling/tile_graph.py:243-257 +
/root/reference/util/data_util.py:167-169: ``[0, len, one-hot(d_e - 2)]`` with ``len`` in
{0.5, 0.866025, 1.0}; the (type, len) pair is a symmetric hash of the endpoints, so there are at
most 3 * (d_e - 2) distinct rows (51 for d_e = 19; the shipped graphs have 20-41).
Node features mimic data_util.py:185-189: one-hot(tile id) ++ area ratio.

Everything is pure integer hashing on torch tensors, so the same graph comes out on CPU and GPU,
and any destination range [lo, hi) can be generated on its own (sharded runs never build the whole
graph on one rank).
"""
from __future__ import annotations

import math

import torch

_LENS = (0.5, 0.8660254037844386, 1.0)


def _stencil(count, skip=0):
    """``count`` lattice offsets closed under negation, ordered by distance, after skipping ``skip``."""
    r = 2 + int(math.ceil(math.sqrt((count + skip) / 2.0)))
    offs = [(dr, dc) for dr in range(-r, r + 1) for dc in range(-r, r + 1)
            if (dr, dc) > (0, 0)]                       # one of each +-pair
    offs.sort(key=lambda o: (o[0] * o[0] + o[1] * o[1], o))
    half = offs[skip // 2: skip // 2 + (count + 1) // 2]
    out = []
    for dr, dc in half:
        out += [(dr, dc), (-dr, -dc)]
    return out[:count] if count % 2 == 0 else out[:count + 1]


def _hash2(a, b, seed):
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    h = (lo * 2654435761) ^ (hi * 2246822519) ^ (seed * 40503 + 12345)
    h = (h ^ (h >> 15)) * 73244475
    return (h ^ (h >> 13)) & 0x7FFFFFFF


def _edges(n, side, lo, hi, stencil, device):
    ids = torch.arange(lo, hi, dtype=torch.int64, device=device)
    r, c = ids // side, ids % side
    srcs, dsts = [], []
    for dr, dc in stencil:
        rr, cc = r + dr, c + dc
        s = rr * side + cc
        ok = (rr >= 0) & (rr < side) & (cc >= 0) & (cc < side) & (s < n)
        srcs.append(s[ok])
        dsts.append(ids[ok])
    if not srcs:
        z = torch.zeros(0, dtype=torch.int64, device=device)
        return z, z.clone()
    return torch.cat(srcs), torch.cat(dsts)


def node_features(n, d_x=3, seed=0, device="cpu", lo=0, hi=None):
    hi = n if hi is None else hi
    ids = torch.arange(lo, hi, dtype=torch.int64, device=device)
    h = _hash2(ids, ids + 7919, seed)
    x = torch.zeros(hi - lo, d_x, dtype=torch.float32, device=device)
    x[torch.arange(hi - lo, device=device), h % (d_x - 1)] = 1.0
    x[:, -1] = 1.0 - 0.5 * ((h >> 8) & 1).to(torch.float32)
    return x


def lattice_graph(n, deg_adj, deg_col, d_x=3, d_e=19, seed=0, device="cpu", lo=0, hi=None,
                  continuous_features=False):
    """Returns ``(x, adj_e_index[2,E_a], adj_e_features[E_a,d_e], col_e_idx[2,E_c])`` restricted to
    destinations in ``[lo, hi)`` (sources are global ids).  int64 indices, fp32 features, exactly the
    dtypes ``to_torch_tensor`` (util/data_util.py:110-117) produces."""
    hi = n if hi is None else hi
    side = int(math.ceil(math.sqrt(n)))
    a_src, a_dst = _edges(n, side, lo, hi, _stencil(deg_adj), device)
    c_src, c_dst = _edges(n, side, lo, hi, _stencil(deg_col, skip=deg_adj + (deg_adj % 2)), device)
    h = _hash2(a_src, a_dst, seed)
    n_onehot = d_e - 2
    feat = torch.zeros(a_src.numel(), d_e, dtype=torch.float32, device=device)
    if a_src.numel() > 0:
        feat[torch.arange(a_src.numel(), device=device), 2 + (h % n_onehot)] = 1.0
        lens = torch.tensor(_LENS, dtype=torch.float32, device=device)
        feat[:, 1] = lens[(h // n_onehot) % 3]
        if continuous_features:                         # exercises the many-types path: no two rows equal
            feat[:, 0] = (h.to(torch.float64) / 2147483648.0).to(torch.float32)
    x = node_features(n, d_x, seed, device, lo, hi)
    return x, torch.stack([a_src, a_dst]), feat, torch.stack([c_src, c_dst])


def random_graph(n, deg_adj, deg_col, d_x=3, d_e=19, seed=0, device="cpu"):
    """Worst-case gather/halo variant: sources uniform in [0, N) (not symmetric; may contain
    duplicates and collision self loops, which GINConv drops)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dst_a = torch.arange(n, dtype=torch.int64).repeat_interleave(deg_adj)
    dst_c = torch.arange(n, dtype=torch.int64).repeat_interleave(deg_col)
    src_a = torch.randint(0, n, (n * deg_adj,), generator=g)
    src_c = torch.randint(0, n, (n * deg_col,), generator=g)
    h = _hash2(src_a, dst_a, seed)
    n_onehot = d_e - 2
    feat = torch.zeros(src_a.numel(), d_e, dtype=torch.float32)
    feat[torch.arange(src_a.numel()), 2 + (h % n_onehot)] = 1.0
    feat[:, 1] = torch.tensor(_LENS, dtype=torch.float32)[(h // n_onehot) % 3]
    x = node_features(n, d_x, seed)
    return (x.to(device), torch.stack([src_a, dst_a]).to(device), feat.to(device),
            torch.stack([src_c, dst_c]).to(device))
