"""world_size-2 (and 3) gloo tests of the node-range sharding plan (host logic of the N>1 path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, deg, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from tilingnn_b200 import shard, synthetic as syn
        bounds = shard.even_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        x, ai, af, ci = syn.lattice_graph(n, deg, deg, seed=1, lo=lo, hi=hi)
        plan = shard.make_plan(n, bounds, ai, ci)
        gid = plan.global_id_of_local_rows()
        # every remapped source points at a local row that mirrors the right global node
        assert torch.equal(gid[plan.adj_src_local], ai[0])
        assert torch.equal(gid[plan.col_src_local], ci[0])
        assert torch.equal(plan.adj_dst_local + lo, ai[1]) and torch.equal(plan.col_dst_local + lo, ci[1])
        assert plan.adj_dst_local.min() >= 0 and plan.adj_dst_local.max() < plan.n_own
        # the send list is made of own rows only, sorted, and equals what this rank publishes
        assert (plan.send_rows >= 0).all() and (plan.send_rows < plan.n_own).all()
        assert torch.equal(plan.send_rows + lo, plan.publish_lists[rank])
        assert all(p.numel() <= plan.halo_slot for p in plan.publish_lists)
        # halo is a boundary effect on a lattice: far smaller than the shard
        assert plan.send_rows.numel() < 0.5 * plan.n_own
        # emulate one halo exchange with gloo and check a gather through local rows
        feat = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, 4)     # row i holds its global id
        local = torch.zeros(plan.n_rows, 4)
        local[: plan.n_own] = feat
        slot = torch.zeros(plan.halo_slot, 4)
        slot[: plan.send_rows.numel()] = feat[plan.send_rows]
        slots = [torch.zeros_like(slot) for _ in range(world)]
        dist.all_gather(slots, slot)
        local[plan.n_own:] = torch.cat(slots)
        assert torch.equal(local[plan.adj_src_local][:, 0], ai[0].float())
        assert torch.equal(local[plan.col_src_local][:, 0], ci[0].float())
        # peers-only exchange (tgnn_set_halo_peers): a row goes only to the ranks whose bit is set in send_mask.  Emulate it
        # (NaN wherever nothing was sent) -- every row the local edges read must still have arrived
        assert plan.send_mask is not None and plan.send_mask.numel() == plan.send_rows.numel()
        assert (plan.send_mask != 0).all() and ((plan.send_mask >> rank) & 1 == 0).all()
        masks = [torch.zeros(plan.halo_slot, dtype=torch.uint8) for _ in range(world)]
        mine = torch.zeros(plan.halo_slot, dtype=torch.uint8)
        mine[: plan.send_mask.numel()] = plan.send_mask
        dist.all_gather(masks, mine)
        local2 = torch.full((plan.n_rows, 4), float("nan"))
        local2[: plan.n_own] = feat
        for qq in range(world):
            got = ((masks[qq] >> rank) & 1).bool()
            base = plan.n_own + qq * plan.halo_slot
            local2[base: base + plan.halo_slot][got] = slots[qq][got]
        assert torch.equal(local2[plan.adj_src_local][:, 0], ai[0].float())
        assert torch.equal(local2[plan.col_src_local][:, 0], ci[0].float())
        if world >= 3 and rank == 0:       # a 1-D lattice shard talks to its neighbours only: rank 0 sends nothing to rank 2
            assert ((plan.send_mask >> 2) & 1 == 0).all()
        q.put((rank, "ok", int(plan.halo_slot), int(plan.n_own)))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world,n,deg", [(2, 6000, 8), (3, 5000, 32)])
def test_shard_plan_gloo(world, n, deg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, deg, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert sum(r[3] for r in res) == n


def test_even_bounds():
    from tilingnn_b200 import shard
    for n, w in ((1000000, 8), (10000, 2), (200, 4), (65, 2)):
        b = shard.even_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        assert all(b[i] < b[i + 1] for i in range(w))                    # no rank is left without nodes
        assert all(x % 64 == 0 or x == n for x in b[:-1])
    for n, w in ((100, 4), (63, 2)):                                     # would leave trailing ranks empty: rejected everywhere
        with pytest.raises(ValueError):
            shard.even_bounds(n, w)


def test_bind_to_gpu_numa_is_optional():
    """``shard.bind_to_gpu_numa`` (NVML CPU affinity of the rank's GPU -> sched_setaffinity) must never get in the way: without a
    GPU / NVML it returns None and leaves the process's affinity mask alone."""
    import os
    from tilingnn_b200 import shard
    before = os.sched_getaffinity(0)
    assert shard.bind_to_gpu_numa(0) is None or isinstance(shard.bind_to_gpu_numa(0), int)
    if not __import__("torch").cuda.is_available():
        assert os.sched_getaffinity(0) == before
