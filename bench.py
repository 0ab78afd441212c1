#!/usr/bin/env python
"""Benchmark of the TilinGNN scoring forward pass (BASELINE.json: node-scores/sec on a 1M-node,
avg-degree-32 super-graph; HBM GB/s vs roofline).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU algorithm (oracle port)

A "step" is one full forward (6 message-passing layers + init/final MLP, train-mode BatchNorm --
the reference's behaviour) over one synthetic super-graph of ``--nodes`` nodes per GPU (weak
scaling: every rank owns ``--nodes`` nodes of an N x larger lattice graph, halo all-gather + BN
all-reduce per layer).  ``value`` is timed with CUDA events with the graph resident in HBM;
``e2e`` goes through the reference-facing call (host numpy arrays in, host scores out, graph
structures rebuilt every call as the reference's per-call topology requires).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "node-scores/sec"
UNIT = "node-scores/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tilingnn", choices=["tilingnn", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000, help="nodes per GPU (weak scaling) / in total (strong scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns --nodes nodes of an N x larger graph; strong: ONE --nodes graph is cut into "
                         "N node-range shards (BASELINE.json config 4)")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-vs-oracle parity check of multi-GPU runs")
    ap.add_argument("--deg", type=int, default=32, help="adjacency and collision stencil size")
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--bn", default="train", choices=["train", "eval"])
    ap.add_argument("--graph", default="lattice", choices=["lattice", "random"])
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-nodes", type=int, default=0, help="0 = calibrate to ~15 s")
    ap.add_argument("--config5", action="store_true",
                    help="BASELINE.json config 5 instead: 30-60-90+equilateral, bunny.txt, 4 layouts, scoring + greedy "
                         "assembly wall-clock (fixtures under tests/golden/), CPU port beside it")
    return ap.parse_args()


D_X, D_E = 3, 19


def bytes_per_node_model(L, d_adj, d_col, bn):
    """ALGORITHMIC bytes per node-score (SURVEY.md §8d / BASELINE.md §3)."""
    if bn == "train":
        return L * (1032 + 8 * d_adj + 4 * d_col) - 256 + (8 * D_X + 128) + (128 * (L + 1) + 3844)
    return L * (648 + 8 * d_adj + 4 * d_col) - 256 + (4 * D_X + 128) + (128 * (L + 1) + 4)


def kernel_bytes_model(fam, n, e_adj, e_col, L):
    """ALGORITHMIC bytes of ONE launch of a kernel family."""
    if fam == "conv":      # read b1 rows once, write pre1, 8 B per adjacency edge (index + type/dst), 4 B/row
        return n * 260 + 8 * e_adj
    if fam == "gin":       # read pre2 rows once, write pre2', 4 B per collision edge, row pointers
        return n * 260 + 4 * e_col
    if fam == "combine":   # pre1, pre2, residual in; b1 out
        return n * 512
    if fam == "final":     # averaged over the four dense stages: 128(L+1) + 2*4*(256+128+64+32) - 128 (a3 read by score)
        return n * (128 * (L + 1) + 8 * 480 - 128) / 4
    if fam == "init":
        return n * (8 * D_X + 128) / 3
    return 0


def kernel_flops_model(fam, n, e_adj, e_col, L):
    """ALGORITHMIC tensor-core FLOPs of ONE launch of a kernel family, times the THREE
    half-precision products a split-precision fp32-accurate contraction costs (hi.hi + hi.lo + lo.hi)."""
    if fam == "conv":      # per adjacency edge a [1x32].[32x32] product, per node the root term
        return 3 * 2.0 * 1024 * (e_adj + n)
    if fam == "gin":       # node MLP 32 -> 32 -> 64 -> 32
        return 3 * 2.0 * 5120 * n
    if fam == "final":     # averaged over the four dense stages
        return 3 * 2.0 * (32 * (L + 1) * 256 + 256 * 128 + 128 * 64 + 64 * 32) * n / 4
    return 0.0


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s, source)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1640.9)), "measured (MEASURED_PEAKS.json, burst figures)"
        except Exception:
            pass
    return 6650.0, 1640.9, "fallback (B200_PROFILING.md)"


def sharded_parity(net_factory, dev, rank, world, group=None):
    """Driver-visible multi-GPU parity: two small lattices scored through the SHARDED path (same exchange kernels as
    the benchmark) and compared on rank 0 with the fp64 oracle of the UNSHARDED graph.  Returns the max abs error."""
    import torch
    import torch.distributed as dist
    from tilingnn_b200 import shard as shard_mod, synthetic as syn
    worst = 0.0
    for n, deg in ((20000, 8), (6000, 32)):
        bounds = shard_mod.even_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        x, ai, af, ci = syn.lattice_graph(n, deg, deg, D_X, D_E, seed=1, device=dev, lo=lo, hi=hi)
        net, params = net_factory()
        plan = shard_mod.make_plan(n, bounds, ai, ci)
        net.set_graph_shard(plan, af)
        s = net.score(x)
        net.check_errors()
        parts = [torch.zeros(bounds[q + 1] - bounds[q], dtype=torch.float32, device=dev) for q in range(world)]
        dist.all_gather(parts, s.contiguous())
        if rank == 0:
            from oracle import tilingnn_oracle as orc          # the checker (fp64, unsharded), never the thing measured
            xg, aig, afg, cig = syn.lattice_graph(n, deg, deg, D_X, D_E, seed=1)
            gold = orc.forward(params, xg, aig, afg, cig, depth=net.network_depth, bn_mode="train", dtype=torch.float64)[:, 0]
            worst = max(worst, float((torch.cat(parts).double().cpu() - gold).abs().max()))
        del net
    t = torch.tensor([worst], device=dev)
    dist.broadcast(t, src=0)
    return float(t.item())


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        for r in rows:
            f = [s.strip() for s in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def seeded_state_dict(depth):
    """Random-init weights of the benchmarked architecture (torch default init, seed 0) under the
    reference's keys -- shared by the CUDA arm and the CPU baseline."""
    import torch
    from tilingnn_b200 import TilinGNN
    torch.manual_seed(0)
    net = TilinGNN(D_E, depth, 32, node_features_dim=D_X)
    return net, {k: v.clone() for k, v in net.state_dict().items()}


def make_graph(args, n_global, lo, hi, device):
    from tilingnn_b200 import synthetic as syn
    if args.graph == "random":
        assert lo == 0 and hi == n_global, "random graphs are single-GPU only"
        return syn.random_graph(n_global, args.deg, args.deg, D_X, D_E, seed=0, device=device)
    return syn.lattice_graph(n_global, args.deg, args.deg, D_X, D_E, seed=0, device=device, lo=lo, hi=hi)


def cpu_reference_rate(args, steps, warmup, sample_nodes=0, budget_s=15.0):
    """The reference's algorithm (per-edge MLP -> [E,1024] weights, scatter-mean, GIN, train-BN) as
    restated in oracle/tilingnn_oracle.py, fp32, torch CPU with all host threads."""
    import torch
    from oracle import tilingnn_oracle as orc
    from tilingnn_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    _, sd = seeded_state_dict(args.depth)

    def one(n):
        x, ai, af, ci = syn.lattice_graph(n, args.deg, args.deg, D_X, D_E, seed=0)
        t = time.perf_counter()
        with torch.no_grad():
            orc.forward(sd, x, ai, af, ci, depth=args.depth, bn_mode=args.bn, dtype=torch.float32)
        return time.perf_counter() - t
    n = sample_nodes
    if n <= 0:
        one(500)                                              # thread-pool / allocator warm-up
        t_cal = one(2000)
        per_step = budget_s / max(1, steps + warmup)
        n = int(min(args.nodes, max(1000, 2000 * per_step / max(t_cal, 1e-3))))
    for _ in range(warmup):
        one(n)
    ts = [one(n) for _ in range(max(1, steps))]
    mean = sum(ts) / len(ts)
    return {"value": n / mean, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n}-node lattice, deg {args.deg}+{args.deg}, depth {args.depth}, fp32, {args.bn}-BN, "
                      f"{len(ts)} step(s) of {mean:.2f} s on {cores} threads (oracle/tilingnn_oracle.py)"}, mean, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, mean, n = cpu_reference_rate(args, args.steps, args.warmup, args.cpu_sample_nodes, budget_s=120.0)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic lattice super-graph, avg-deg {args.deg} adj + {args.deg} col, "
                                   f"{args.depth} layers, width 32, {args.bn}-mode BatchNorm; CPU sample of {n} nodes "
                                   f"(the reference materialises [E,1024] fp32 per layer: 131 GB at 1M nodes)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_config5(args):
    """BASELINE.json config 5: the four bunny layouts of Tiling-Shape.py:52-54 (30-60-90+equilateral, shipped
    checkpoint, depth 20, train-mode BatchNorm) solved by ``ML_Solver.solve`` = greedy assembly with every round
    scored by the CUDA network; beside it the CPU port (oracle network, fp32, all host threads) driving the SAME
    greedy loop -- the cpu_baseline leg, the only place this file touches oracle/.  Prints ONE JSON line."""
    import numpy as np
    import torch
    import __graft_entry__ as ge
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _util import GOLDEN, load_ckpt, load_layout
    from tilingnn_b200 import ML_Solver, TilinGNN, greedy
    ge.build()
    z = dict(np.load(os.path.join(GOLDEN, "c5_bunny.npz")))
    ckpt = load_ckpt("ckpt_30-60-90+equilateral.npz")
    layouts = [load_layout(z, prefix=f"L{i}_") for i in range(int(z["n_layouts"]))]
    dev = torch.device("cuda:0")
    net = TilinGNN(int(z["d_e"]), 20, 32, node_features_dim=int(z["d_x"]))
    net.load_state_dict(ckpt, strict=True)
    net = net.to(dev).train()
    solver = ML_Solver(None, dev, None, net, 1)             # one solver for all layouts, as Tiling-Shape.py:37
    calls = {"n": 0}

    def run_gpu(seed):
        out, rng = [], np.random.RandomState(seed)
        for sg, graph in layouts:
            solver.complete_graph = graph
            solved, score = solver.solve(sg, rng=rng)
            calls["n"] += solved.greedy_rounds + 1
            out.append((int(solved.predict.sum()), score))
        return out
    for _ in range(max(1, args.warmup // 3)):
        run_gpu(0)                                          # library load, parameter upload, cached contour areas
    torch.cuda.synchronize()
    times = []
    for _ in range(max(1, args.steps // 4)):
        calls["n"] = 0
        t0 = time.perf_counter()
        res = run_gpu(2)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    gpu_s = float(np.median(times))

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import tilingnn_oracle as orc           # cpu_baseline leg

        class OracleSolver:
            def __init__(self, graph):
                self.complete_graph, self.calls = graph, 0

            def predict(self, lay):
                n = lay.node_feature.shape[0]
                if np.size(lay.collide_edge_index) == 0 or np.size(lay.align_edge_index) == 0:
                    return np.ones(n, dtype=np.float32)
                self.calls += 1
                t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
                s = orc.forward(ckpt, t(lay.node_feature, torch.float32), t(lay.align_edge_index, torch.long),
                                t(lay.align_edge_features, torch.float32), t(lay.collide_edge_index, torch.long), depth=20,
                                bn_mode="train", dtype=torch.float32)
                return s[:, 0].float().numpy()
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        rng, cpu_res, cpu_calls = np.random.RandomState(2), [], 0
        for sg, graph in layouts:
            s = OracleSolver(graph)
            r = greedy.solve_by_probablistic_greedy(s, sg, rng=rng)
            s.predict(sg)                                   # ML_Solver.solve's final scoring pass (ml_solver.py:65)
            cpu_calls += s.calls
            cpu_res.append((int(r.selection.sum()), r.score))
        cpu_s = time.perf_counter() - t0
        cpu = {"value": cpu_s, "unit": "s", "cores": os.cpu_count(), "kind": "port", "network_calls": cpu_calls,
               "sample": "the same 4 layouts, oracle network fp32 driving the same greedy loop (one run)",
               "layouts": [{"tiles_placed": a, "score": b} for a, b in cpu_res]}
    print(json.dumps({
        "metric": "Tiling-Shape scoring + greedy assembly wall-clock (config 5)", "unit": "s", "higher_is_better": False,
        "value": gpu_s, "times": times, "network_calls": calls["n"], "n_gpus": 1, "dtype": "f32",
        "data": "tests/golden/c5_bunny.npz + shipped 30-60-90+equilateral checkpoint",
        "layouts": [{"nodes": int(sg.node_feature.shape[0]), "tiles_placed": a, "score": b} for (sg, _), (a, b) in zip(layouts, res)],
        "cpu_baseline": cpu, "speedup_vs_cpu_port": (cpu["value"] / gpu_s) if cpu else None,
        "config": {"workload": "30-60-90+equilateral, bunny.txt, 4 layouts (604/562/591/565 candidate tiles), depth 20, "
                               "train-mode BatchNorm, shipped checkpoint"}}))


def main():
    args = parse()
    if args.config5:
        return run_config5(args)
    if args.impl == "reference":
        return run_reference(args)
    # NCCL writes its banner to stdout when NCCL_DEBUG=VERSION/INFO is set in the environment; stdout must carry
    # exactly one JSON line.  TGNN_NCCL_DEBUG passes a level through explicitly (output then goes to stderr's file).
    # NCCL writes its log to stdout by default; stdout must carry exactly one JSON line, so the log goes to stderr --
    # at whatever level the caller asked for (NCCL_DEBUG is left alone: the driver reads the communicator sizes from it).
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0 and not os.path.exists(ge.LIB):
        ge.build()                                   # normally prebuilt in-tree and shipped with the snapshot
    if world > 1:
        dist.barrier()
    from tilingnn_b200 import shard as shard_mod
    # one process per GPU: keep each rank (and the pinned buffers it allocates) on the NUMA node of its GPU
    numa_cores = shard_mod.bind_to_gpu_numa(local) if world > 1 and os.environ.get("TGNN_NUMA_BIND", "1") != "0" else None
    warmup = max(3, args.warmup)

    n_global = args.nodes * world if args.scaling == "weak" else args.nodes
    bounds = shard_mod.even_bounds(n_global, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    net, sd = seeded_state_dict(args.depth)
    net = net.to(dev)
    net.train() if args.bn == "train" else net.eval()
    parity_err = None
    if world > 1 and not args.no_parity:
        def small_net():
            from oracle import tilingnn_oracle as orc
            from tilingnn_b200 import TilinGNN
            p = orc.make_params(D_X, D_E, args.depth, seed=1)          # conditioned weights, identical on every rank
            m = TilinGNN(D_E, args.depth, 32, node_features_dim=D_X)
            m.load_state_dict(p, strict=True)
            m = m.to(dev).train()
            m.shard_init()
            return m, p
        parity_err = sharded_parity(small_net, dev, rank, world)
        if not (parity_err <= 1e-4):
            raise SystemExit(f"sharded parity check failed: max |sharded CUDA - unsharded fp64 oracle| = {parity_err:.3e} > 1e-4")
    x, ai, af, ci = make_graph(args, n_global, lo, hi, dev)
    e_adj, e_col, n_own = ai.shape[1], ci.shape[1], hi - lo
    plan = None
    if world > 1:
        net.shard_init()
        plan = shard_mod.make_plan(n_global, bounds, ai, ci)
        net.set_graph_shard(plan, af)
    else:
        net.set_graph(n_own, ai, af, ci)
    out = torch.empty(n_own, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        net.score(x, out=out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        net.score(x, out=out)
    ev1.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_global / (ms_step * 1e-3)
    info = net.info()
    assert bool(torch.isfinite(out).all()), "non-finite scores"

    # ---- per-kernel-family device times (CUDA events on the launch stream, inside the library) ----
    net.set_profiling(True)
    fam_ms = {}
    reps = 3
    for _ in range(reps):
        net.score(x, out=out)
        for fam, (ms, nl) in net.profile().items():
            a = fam_ms.setdefault(fam, [0.0, 0])
            a[0] += ms / reps; a[1] = nl
    net.set_profiling(False)
    torch.cuda.synchronize()
    peak, peak_tc, peak_src = measured_peaks()
    dom = max(fam_ms, key=lambda f: fam_ms[f][0])
    dom_ms, dom_launches = fam_ms[dom]
    per_launch_ms = dom_ms / max(1, dom_launches)
    kb = kernel_bytes_model(dom, n_own, e_adj, e_col, args.depth)
    kf = kernel_flops_model(dom, n_own, e_adj, e_col, args.depth)
    achieved = kb / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    achieved_tf = kf / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
    t_hbm_ms, t_tc_ms = kb / (peak * 1e9) * 1e3, kf / (peak_tc * 1e12) * 1e3
    bound = "tensor" if t_tc_ms > t_hbm_ms else "hbm"
    traffic, traffic_src = None, None
    prof_json = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof_json):
        try:
            pj = json.load(open(prof_json))
            traffic = pj.get("traffic_bytes_per_launch", {}).get(dom)
            traffic_src = (f"profiles/ncu_summary.json: dram__bytes_read+write of ONE ncu --set full capture "
                           f"({pj.get('captured_on', 'single GPU, headline workload')}); not re-measured in this run")
        except Exception:
            traffic = None
    bpn = bytes_per_node_model(args.depth, e_adj / n_own, e_col / n_own, args.bn)
    fwd_gbs = bpn * n_own / (ms_step * 1e-3) / 1e9

    # ---- end to end through the reference-facing call, HOST buffers in / out -----------------------
    e2e = None
    if not args.no_e2e:
        h = [t.cpu().pin_memory() for t in (x, ai, af, ci)]
        h2d = sum(t.numel() * t.element_size() for t in h)
        host_out = torch.empty(n_own, dtype=torch.float32).pin_memory()
        del ai, af, ci
        torch.cuda.empty_cache()

        # Streamed: while graph k is built and scored, the arrays of graph k+1 travel host -> device on a copy stream
        # (tilingnn_b200.streaming.ScoreStream, the public call for a sequence of layouts).  Every step copies ITS inputs
        # from pinned host memory and reads ITS scores back; `latency_ms` is the same step run alone (nothing overlapped).
        from tilingnn_b200.streaming import ScoreStream

        def dev_step(d):
            if world > 1:
                p = shard_mod.make_plan(n_global, bounds, d[1], d[3])
                net.set_graph_shard(p, d[2])
                return net.score(d[0])
            return net(x=d[0], adj_e_index=d[1], adj_e_features=d[2], col_e_idx=d[3])[0][:, 0]
        stream = ScoreStream(net, dev_step)
        stream([h], [host_out])                      # warm-up (allocates the device slots)
        torch.cuda.synchronize()
        t_l = time.perf_counter()
        stream([h], [host_out])                      # one step alone: copy -> build -> forward -> read back
        torch.cuda.synchronize()
        latency_ms = (time.perf_counter() - t_l) * 1e3
        barrier()
        t_a = time.perf_counter()
        stream([h] * args.e2e_steps, [host_out] * args.e2e_steps)
        torch.cuda.synchronize()
        barrier()
        dt = (time.perf_counter() - t_a) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n_global / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(n_own * 4),
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "latency_ms": latency_ms,
               "note": "per step: pinned host arrays -> H2D -> graph structures rebuilt -> forward -> D2H scores; steps are streamed "
                       "(double-buffered device slots: the H2D copy of step k+1 overlaps the build + forward of step k); "
                       "latency_ms = one step alone; per rank bytes"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_reference_rate(args, steps=1, warmup=0, sample_nodes=args.cpu_sample_nodes)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic {args.graph} super-graph, {n_own} nodes per GPU ({n_global} total, {args.scaling} scaling), "
                                   f"avg-deg {e_adj / n_own:.1f} adj + {e_col / n_own:.1f} col, {args.depth} layers, width 32, "
                                   f"{args.bn}-mode BatchNorm (reference behaviour), {info['n_edge_types']} edge types",
                       "nodes_per_gpu": n_own, "deg": args.deg, "depth": args.depth, "bn": args.bn,
                       "parallelism": f"node-range shards x{world}" if world > 1 else "single GPU",
                       "l2_policy": f"no flush needed: per-step working set {info['workspace_bytes'] / 1e9:.1f} GB >> 126 MB L2"},
            "roofline": {"bound": bound, "kernel": {"conv": {0: "k_conv_adj", 1: "k_conv_s", 2: "k_conv_h", 3: "k_conv_t", 4: "k_conv_z", 5: "k_conv_x"}[int(info["conv_kernel"])],
                                                    "gin": "k_gin_w" if int(info.get("gin_kernel", 0)) else "k_gin", "final": "k_dense_tc", "combine": "k_combine"}.get(dom, dom),
                         "achieved": achieved_tf if bound == "tensor" else achieved,
                         "peak": peak_tc if bound == "tensor" else peak,
                         "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                         "frac": (achieved_tf / peak_tc) if bound == "tensor" else (achieved / peak),
                         "t_hbm_ms": t_hbm_ms, "t_tc_ms": t_tc_ms,
                         "hbm": {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak},
                         "tensor": {"achieved": achieved_tf, "peak": peak_tc, "unit": "TFLOP/s", "frac": achieved_tf / peak_tc,
                                    "algorithmic_flops_per_launch": kf,
                                    "note": "3 half-precision products per fp32-accurate contraction (split precision)"},
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kb, "ms_per_launch": per_launch_ms,
                         "share_of_step": dom_ms / max(1e-9, sum(v[0] for v in fam_ms.values())),
                         "forward": {"algorithmic_bytes_per_node": bpn, "achieved": fwd_gbs, "frac": fwd_gbs / peak}},
            "kernel_ms": {f: round(v[0], 4) for f, v in fam_ms.items()},
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(info["launches_per_forward"]) * args.steps,
            "collectives_per_step": int(info["collectives_per_forward"]),
            "parity_max_err": parity_err,
            "numa_bound_cores": numa_cores,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
