"""Node-range sharding plan (host side, backend agnostic: gloo on CPU in the tests, NCCL on GPUs).

The reference is single-device; this is the new multi-GPU layer of SURVEY.md §8e.  Rank r owns the
contiguous node range ``[bounds[r], bounds[r+1])`` -- rows of both CSRs, features and scores of
those nodes.  Source rows owned by peers ("halo") are mirrored once per layer by ONE all-gather of
the rows every rank publishes (its *send list*: own rows that at least one peer reads).

Local row numbering used by libtgnn (include/tgnn.h, ``tgnn_set_graph_shard``):
    [0, n_own)                                   own rows, global id = lo + i
    [n_own + q*halo_slot, n_own + (q+1)*halo_slot)   the send list of rank q (padded to halo_slot)
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


def even_bounds(n_global: int, world: int):
    """Contiguous, 64-aligned (warp-tile aligned) node ranges of near-equal size."""
    per = -(-n_global // world)
    per = -(-per // 64) * 64
    bounds = [min(n_global, r * per) for r in range(world)] + [n_global]
    if any(b1 <= b0 for b0, b1 in zip(bounds, bounds[1:])):
        # every rank computes the same bounds, so every rank raises here -- before any collective is entered
        # (an empty rank would otherwise fail alone in tgnn_set_graph_shard and leave its peers blocked in NCCL)
        raise ValueError(f"even_bounds: {n_global} nodes cannot be cut into {world} non-empty 64-aligned ranges "
                         f"(need more than {64 * (world - 1)} nodes)")
    return bounds


@dataclass
class ShardPlan:
    rank: int
    world: int
    lo: int
    hi: int
    n_global: int
    halo_slot: int
    send_rows: torch.Tensor        # int64 [n_send] local row ids (sorted) this rank publishes
    adj_src_local: torch.Tensor    # int64 [E_a] remapped sources
    adj_dst_local: torch.Tensor    # int64 [E_a] dst - lo
    col_src_local: torch.Tensor
    col_dst_local: torch.Tensor
    publish_lists: list            # per rank: int64 global ids it publishes (sorted)
    send_mask: torch.Tensor = None # uint8 [n_send]: bit q = rank q reads send row i (world <= 8; else None = every peer)

    @property
    def n_own(self):
        return self.hi - self.lo

    @property
    def n_rows(self):
        return self.n_own + self.world * self.halo_slot

    def global_id_of_local_rows(self):
        """int64 [n_rows]: global node id behind every local row (-1 for padding)."""
        out = torch.full((self.n_rows,), -1, dtype=torch.int64)
        out[: self.n_own] = torch.arange(self.lo, self.hi)
        for q, ids in enumerate(self.publish_lists):
            base = self.n_own + q * self.halo_slot
            out[base: base + ids.numel()] = ids.cpu()
        return out


def _all_gather_var(t: torch.Tensor, group):
    """all_gather of 1-D int64 tensors of different lengths (pad to the max length)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    pad = torch.full((m,), -1, dtype=torch.int64, device=t.device)
    pad[: t.numel()] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:c] for o, c in zip(outs, counts)]


def make_plan(n_global, bounds, adj_index, col_index, group=None) -> ShardPlan:
    """``adj_index`` / ``col_index``: int64 [2, E] with GLOBAL ids whose destinations all lie in this
    rank's range.  Collective: every rank of ``group`` must call it."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    dev = adj_index.device
    for idx in (adj_index, col_index):
        if idx.numel() and (int(idx[1].min()) < lo or int(idx[1].max()) >= hi):
            raise ValueError("make_plan: every destination must be owned by this rank")
    srcs = torch.cat([adj_index[0].reshape(-1), col_index[0].reshape(-1)])
    remote = srcs[(srcs < lo) | (srcs >= hi)]
    need = torch.unique(remote)                                   # sorted global ids this rank reads from peers
    needs = _all_gather_var(need, group)
    mine = [nq[(nq >= lo) & (nq < hi)] for q, nq in enumerate(needs) if q != rank]
    publish = torch.unique(torch.cat(mine)) if mine else torch.zeros(0, dtype=torch.int64, device=dev)
    publish_lists = _all_gather_var(publish, group)
    halo_slot = max(max(p.numel() for p in publish_lists), 1)
    # which peer reads which of my published rows: a boundary row then travels only to those ranks (tgnn_set_halo_peers)
    send_mask = None
    if world <= 8:
        send_mask = torch.zeros(publish.numel(), dtype=torch.uint8, device=dev)
        for q, nq in enumerate(needs):
            if q == rank:
                continue
            ids = nq[(nq >= lo) & (nq < hi)]
            if ids.numel():
                send_mask[torch.searchsorted(publish, ids)] |= (1 << q)
    n_own = hi - lo
    b = torch.tensor(bounds, dtype=torch.int64, device=dev)

    def remap(src):
        out = src - lo
        rem = (src < lo) | (src >= hi)
        if rem.any():
            ids = src[rem]
            owner = torch.bucketize(ids, b, right=True) - 1
            loc = torch.empty_like(ids)
            for q in range(world):
                sel = owner == q
                if sel.any():
                    pos = torch.searchsorted(publish_lists[q], ids[sel])
                    if not torch.equal(publish_lists[q][pos], ids[sel]):
                        raise RuntimeError("make_plan: a needed row is missing from its owner's publish list")
                    loc[sel] = n_own + q * halo_slot + pos
            out = out.clone()
            out[rem] = loc
        return out

    return ShardPlan(rank, world, lo, hi, int(n_global), int(halo_slot), (publish - lo).contiguous(),
                     remap(adj_index[0]).contiguous(), (adj_index[1] - lo).contiguous(),
                     remap(col_index[0]).contiguous(), (col_index[1] - lo).contiguous(), publish_lists, send_mask)


def bind_to_gpu_numa(device_index):
    """One process per GPU: restrict this process to the CPU cores next to its GPU (NVML's CPU affinity of the device), so
    that the pinned host buffers it allocates afterwards are placed on that NUMA node.  Eight ranks that each stream
    3.5 GB per step out of ONE socket's memory share that socket's DRAM and the inter-socket link (round 1: end-to-end
    efficiency 0.47 at 8 GPUs).  Returns the number of cores kept, or None when NVML / the affinity call is not available
    (the caller carries on unbound)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = torch.cuda.get_device_properties(device_index).uuid
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
            cpus &= os.sched_getaffinity(0)
            if not cpus:
                return None
            os.sched_setaffinity(0, cpus)
            return len(cpus)
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None
