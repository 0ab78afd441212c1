"""Generate the golden vectors under tests/golden/ (run in the build container only).

    python tests/golden/make_golden.py

Needs /root/reference.  Every expected output here is produced by the
reference's OWN graph_networks modules (imported unmodified through
oracle/ref_harness.py), NOT by oracle/tilingnn_oracle.py -- these files are
what pins the oracle (tests/test_oracle_golden.py) and, on the GPU box where
/root/reference does not exist, what the CUDA path is compared with
(tests/test_gpu_parity.py).

Files written
  ckpt_30-60-90.npz     the shipped checkpoint pre-trained_models/30-60-90.pth, fp32, the
                        aliased nnConv.nn.mlp.* keys dropped (recreated on load)
  c1_heart.npz          config 1: heart.txt cropped as Tiling-Shape.py:52-54 does (first layout)
  c1_complete.npz       the full 30-60-90 complete graph (N = 3719)
  syn_small.npz         a seeded default-init reference network (depth 3) on a ragged random
                        graph: isolated nodes, self loops, duplicate edges, continuous edge features
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                      # noqa: E402
from tilingnn_b200 import tile_graph_io as tio            # noqa: E402

REF = "/root/reference"


def pack_graph(x, ai, af, ci):
    """Compact encoding: adjacency features as unique rows + per-edge row id."""
    uniq, inv = np.unique(np.asarray(af, dtype=np.float32), axis=0, return_inverse=True)
    return dict(x=np.asarray(x, dtype=np.float32),
                adj_index=np.asarray(ai, dtype=np.int32),
                adj_feat_rows=uniq.astype(np.float32),
                adj_feat_id=inv.reshape(-1).astype(np.int16),
                col_index=np.asarray(ci, dtype=np.int32))


def ref_outputs(sd, x, ai, af, ci, depth):
    x, af = torch.as_tensor(x).float(), torch.as_tensor(af).float()
    ai, ci = torch.as_tensor(ai).long(), torch.as_tensor(ci).long()
    out = {}
    for mode in ("train", "eval"):
        for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
            s = rh.run_reference(sd, x, ai, af, ci, depth=depth, bn_mode=mode, dtype=dt)
            out[f"ref_{mode}_{name}"] = s[:, 0].double().numpy()
    return out


def ref_intermediates(sd, x, ai, af, ci, depth, layers, mode="train"):
    """Per-layer tensors of the reference module (fp64) captured with forward hooks."""
    d_x, d_e = x.shape[1], af.shape[1]
    net = rh.reference_network(d_x, d_e, depth)
    net.load_state_dict(sd, strict=True)
    net = net.double()
    net.train() if mode == "train" else net.eval()
    cap = {}

    def hook(tag):
        def f(_m, _i, o):
            cap[tag] = (o[0] if isinstance(o, tuple) else o).detach().numpy().copy()
        return f
    net.init_node_feature_trans.register_forward_hook(hook("h0"))
    for i in layers:
        net.brch_1_graph_conv_layers[i].register_forward_hook(hook(f"g1_{i}"))
        net.brch_2_coll_conv_layers[i].register_forward_hook(hook(f"g2_{i}"))
    with torch.no_grad():
        net(x=torch.as_tensor(x).double(), adj_e_index=torch.as_tensor(ai).long(),
            adj_e_features=torch.as_tensor(af).double(), col_e_idx=torch.as_tensor(ci).long())
    return {f"{mode}_{k}": v for k, v in cap.items()}


def main():
    sd = torch.load(f"{REF}/pre-trained_models/30-60-90.pth", map_location="cpu", weights_only=True)
    ck = {k: v.numpy() for k, v in sd.items() if ".nnConv.nn.mlp." not in k}
    np.savez_compressed(f"{HERE}/ckpt_30-60-90.npz", **ck)

    g = tio.load_complete_graph(f"{REF}/data/30-60-90/complete_graph_ring9.pkl", tile_type_count=2)
    ext, ints = tio.load_polygons(f"{REF}/silhouette/heart.txt")
    crops = tio.crop_multiple_layouts_from_contour(ext, ints, g, start_angle=0, end_angle=30, num_of_angle=1,
                                                   movement_delta_ratio=[0, 0.5], margin_padding_ratios=[0.5])
    sizes = np.asarray([[c.node_feature.shape[0], c.align_edge_index.shape[1], c.collide_edge_index.shape[1]]
                        for c in crops])
    c = crops[0]
    d = pack_graph(c.node_feature, c.align_edge_index, c.align_edge_features, c.collide_edge_index)
    d.update(ref_outputs(sd, c.node_feature, c.align_edge_index, c.align_edge_features, c.collide_edge_index, 20))
    d.update(ref_intermediates(sd, c.node_feature, c.align_edge_index, c.align_edge_features,
                               c.collide_edge_index, 20, layers=(0, 1, 5, 19)))
    d["crop_sizes"] = sizes
    d["tiles"] = c.tiles.astype(np.int32)
    np.savez_compressed(f"{HERE}/c1_heart.npz", **d)
    print("c1_heart", sizes.tolist())

    c = tio.complete_super_graph(g)
    d = pack_graph(c.node_feature, c.align_edge_index, c.align_edge_features, c.collide_edge_index)
    d.update(ref_outputs(sd, c.node_feature, c.align_edge_index, c.align_edge_features, c.collide_edge_index, 20))
    np.savez_compressed(f"{HERE}/c1_complete.npz", **d)
    print("c1_complete", c.node_feature.shape, c.align_edge_index.shape, c.collide_edge_index.shape)

    # ragged synthetic case on a default-initialised reference network
    torch.manual_seed(1234)
    rng = np.random.default_rng(1234)
    n, d_x, d_e, depth = 150, 4, 7, 3
    net = rh.reference_network(d_x, d_e, depth)
    for m in net.modules():                                   # make eval-mode BN non-trivial too
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    ssd = {k: v.clone() for k, v in net.state_dict().items()}
    x = rng.random((n, d_x)).astype(np.float32)
    ea, ec = 900, 1100
    ai = rng.integers(0, n - 10, size=(2, ea))                # nodes n-10.. have no adjacency in-edges
    ai[:, :40] = ai[:, 40:80]                                 # duplicate edges
    af = rng.random((ea, d_e)).astype(np.float32)             # continuous features: no two rows equal
    af[:40] = af[40:80]
    ci = rng.integers(5, n, size=(2, ec))                     # nodes 0..4 have no collision edges
    ci[1, :30] = ci[0, :30]                                   # self loops (GINConv removes them)
    d = dict(x=x, adj_index=ai.astype(np.int32), adj_feat=af, col_index=ci.astype(np.int32),
             depth=np.int32(depth))
    d.update(ref_outputs(ssd, x, ai, af, ci, depth))
    d.update(ref_intermediates(ssd, x, ai, af, ci, depth, layers=(0, 1, 2)))
    for k, v in ssd.items():
        if ".nnConv.nn.mlp." not in k:
            d["param:" + k] = v.numpy()
    np.savez_compressed(f"{HERE}/syn_small.npz", **d)
    print("syn_small done")


if __name__ == "__main__":
    main()
