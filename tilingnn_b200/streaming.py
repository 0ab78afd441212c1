"""End-to-end scoring of a STREAM of host-resident super-graphs (the reference's call pattern:
``to_torch_tensor`` -> ``network(...)`` -> ``.cpu()`` once per layout, util/data_util.py:110-117 and
solver/ml_solver/ml_solver.py:35-47), double buffered: while graph k is built and scored on the compute
stream, the arrays of graph k+1 travel host -> device on a copy stream.

At benchmark sizes one step moves 3.46 GB over PCIe (~63 ms) and computes for ~21 ms (structure build +
forward), so the serial call is copy-bound at ~84 ms per graph and the streamed one at ~63 ms: every step's
inputs are still copied from pinned host memory and its scores read back to the host -- only the waiting is
overlapped.  (A single graph has nothing to overlap: its build needs all arrays.)
"""
from __future__ import annotations

import torch


class ScoreStream:
    """``ScoreStream(net)(batches, outs)``: ``batches`` = sequence of ``(x, adj_e_index, adj_e_features, col_e_idx)``
    host tensors (pinned for asynchronous copies), ``outs`` = host tensors ``[N]`` fp32 receiving the scores.
    ``step(device_tensors) -> scores[N]`` replaces the default ``net(...)`` call (sharded runs build their plan there).
    Two device slots are kept and reused; a slot is overwritten only after the forward that read it has finished."""

    def __init__(self, net, step=None):
        self.net = net
        self.dev = net._device()
        if self.dev.type != "cuda":
            raise RuntimeError("ScoreStream needs the network on a CUDA device (no CPU fallback)")
        self.step = step or (lambda d: net(x=d[0], adj_e_index=d[1], adj_e_features=d[2], col_e_idx=d[3])[0][:, 0])
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]

    def _slot(self, i, batch):
        cur = self.slots[i]
        if cur is None or any(c.shape != b.shape or c.dtype != b.dtype for c, b in zip(cur, batch)):
            cur = self.slots[i] = [torch.empty(b.shape, dtype=b.dtype, device=self.dev) for b in batch]
            self.free[i].record(torch.cuda.current_stream(self.dev))
        return cur

    def _upload(self, k, batch):
        i = k & 1
        dst = self._slot(i, batch)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[i])            # the forward that used this slot is done
            for d, b in zip(dst, batch):
                d.copy_(b, non_blocking=True)
            self.ready[i].record(self.copy_stream)

    def __call__(self, batches, outs):
        batches = list(batches)
        if len(outs) != len(batches):
            raise ValueError("one output tensor per batch")
        if not batches:
            return outs
        cur = torch.cuda.current_stream(self.dev)
        self._upload(0, batches[0])
        for k in range(len(batches)):
            if k + 1 < len(batches):
                self._upload(k + 1, batches[k + 1])              # queued before this step's (host-blocking) structure build
            i = k & 1
            cur.wait_event(self.ready[i])
            s = self.step(self.slots[i])
            outs[k].copy_(s, non_blocking=True)
            self.free[i].record(cur)
        return outs


def score_stream(net, batches, outs, step=None):
    """Functional form of :class:`ScoreStream` for one-off use."""
    return ScoreStream(net, step)(batches, outs)
