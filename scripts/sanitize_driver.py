#!/usr/bin/env python
"""Two forwards of the scoring path on a synthetic lattice (single GPU, or node-range shards under torchrun) --
the workload scripts/gpu_sanitize.sh runs under compute-sanitizer.  Checks nothing by itself beyond finiteness."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=10000)
    ap.add_argument("--deg", type=int, default=8)
    ap.add_argument("--depth", type=int, default=6)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tilingnn_b200 import TilinGNN, shard, synthetic as syn
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = TilinGNN(19, a.depth, 32, node_features_dim=3).to(dev).train()
    bounds = shard.even_bounds(a.nodes, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    x, ai, af, ci = syn.lattice_graph(a.nodes, a.deg, a.deg, 3, 19, seed=0, device=dev, lo=lo, hi=hi)
    if world > 1:
        net.shard_init()
        net.set_graph_shard(shard.make_plan(a.nodes, bounds, ai, ci), af)
    else:
        net.set_graph(hi - lo, ai, af, ci)
    for _ in range(2):
        s = net.score(x)
    net.check_errors()
    assert bool(torch.isfinite(s).all())
    print(f"rank {rank}: ok, {hi - lo} nodes, mean score {float(s.mean()):.6f}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
