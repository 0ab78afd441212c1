"""Sharded (node-range) forward on 2 GPUs through NCCL vs the unsharded fp64 oracle.
Needs >= 2 CUDA devices:  gpurun --gpus 2 -- python -m pytest tests/test_gpu_shard.py -m gpu"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, deg, depth, mode, q):
    try:
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from oracle import tilingnn_oracle as orc
        from tilingnn_b200 import TilinGNN, shard, synthetic as syn
        bounds = shard.even_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        x, ai, af, ci = syn.lattice_graph(n, deg, deg, seed=0, device=dev, lo=lo, hi=hi)
        p = orc.make_params(3, 19, depth, seed=0)
        net = TilinGNN(19, depth, 32, node_features_dim=3)
        net.load_state_dict(p)
        net = net.to(dev)
        net.train() if mode == "train" else net.eval()
        net.shard_init()
        plan = shard.make_plan(n, bounds, ai, ci)
        net.set_graph_shard(plan, af)
        s1 = net.score(x).clone()
        s2 = net.score(x).clone()          # second pass: halo buffers are reused
        torch.cuda.synchronize()
        info = net.info()
        q.put((rank, "ok", lo, hi, s1.cpu().numpy(), bool(torch.equal(s1, s2)), info["collectives_per_forward"], plan.halo_slot,
               info["peer_exchange"]))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0, None, False, 0, 0, 0))


@pytest.mark.parametrize("mode", ["train"])
@pytest.mark.parametrize("n,deg,conv,p2p,ginw", [(20000, 8, "h", 1, 0), (6000, 32, "s", 1, 1), (6000, 32, "chunk", 0, 0),
                                                 (20000, 8, "h", 0, 1), (6000, 32, "t", 1, 1)])
def test_two_gpu_shards_match_unsharded_oracle(built_lib, n, deg, conv, p2p, ginw, mode, monkeypatch):
    """both exchange paths: peer-memory stores over NVLink (CUDA IPC, default) and NCCL collectives (TGNN_P2P=0); both
    collision kernels (ginw = 1: staged windows, whose runs then include mirrored halo rows)"""
    monkeypatch.setenv("TGNN_CONV", conv)          # inherited by the spawned ranks
    monkeypatch.setenv("TGNN_P2P", str(p2p))
    monkeypatch.setenv("TGNN_GINW", str(ginw))
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import tilingnn_oracle as orc
    from tilingnn_b200 import synthetic as syn
    depth, world = 6, 2
    x, ai, af, ci = syn.lattice_graph(n, deg, deg, seed=0)
    p = orc.make_params(3, 19, depth, seed=0)
    gold = orc.forward(p, x, ai, af, ci, depth=depth, bn_mode=mode, dtype=torch.float64)[:, 0].numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, deg, depth, mode, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in procs]
    for pr in procs:
        pr.join(timeout=120)
    assert all(r[1] == "ok" for r in res), [r[1] for r in res]
    out = np.zeros(n)
    for rank, _, lo, hi, s, same, ncoll, slot, peer in res:
        out[lo:hi] = s
        assert peer <= p2p, "TGNN_P2P=0 must keep the NCCL collectives"
        if peer != p2p:                                # CUDA IPC refused (container policy): NCCL fallback, still correct
            import warnings
            warnings.warn("peer-memory exchange could not be set up on this box; the NCCL path was tested instead")
        assert same, "sharded forward must be run-to-run deterministic"
        assert ncoll == (2 + depth + 4) + depth, ncoll        # BN all-reduces + halo all-gathers (train mode)
        assert 0 < slot < n // 4
    err = np.abs(out - gold).max()
    print(f"2-GPU sharded N={n} deg={deg} {mode} conv={conv} p2p={p2p} ginw={ginw}: max err vs fp64 oracle {err:.2e}")
    assert err <= 1e-4


def test_two_devices_in_one_process():
    """kernel attributes (opt-in shared memory sizes) are per device: two handles on two GPUs of the same process must
    both work and agree bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import __graft_entry__ as ge
    ge.build()
    from oracle import tilingnn_oracle as orc
    from tilingnn_b200 import TilinGNN, synthetic as syn
    x, ai, af, ci = syn.lattice_graph(5000, 8, 8, seed=1)
    p = orc.make_params(3, 19, 3, seed=1)
    outs = []
    for d in (1, 0):                                   # the second device first
        dev = torch.device("cuda", d)
        net = TilinGNN(19, 3, 32, node_features_dim=3)
        net.load_state_dict(p)
        net = net.to(dev).train()
        s, _ = net(x=x.to(dev), adj_e_index=ai.to(dev), adj_e_features=af.to(dev), col_e_idx=ci.to(dev))
        torch.cuda.synchronize(dev)
        outs.append(s.cpu())
    assert torch.equal(outs[0], outs[1])
