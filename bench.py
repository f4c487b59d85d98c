#!/usr/bin/env python
"""bench.py -- CartNet hot-path throughput (BASELINE.json: crystal graphs/s & edges/s, fwd+bwd, ADP shape) on N B200s
of one node, with the reference CPU path timed beside it.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1      # the reference algorithm on the host cores
    python bench.py --workload jarvis_infer | mp_train | supercell     # BASELINE configs[2..4]

Default workload (config.workload) = BASELINE configs[1], "CartNet ADP training step, batch 64 crystals, 1xB200":
64 synthetic ADP-shaped crystals per GPU (lognormal sizes, mean ~194 atoms, 9.5 A^3/atom, radius 5 A), random-init
CartNet(256, 64, 4 layers, Cholesky head), one step = forward + L1 loss + backward + gradient all-reduce (N > 1) + Adam.
Weak scaling: the global batch is 64 x N crystals, sharded as whole crystals per rank balanced by edge count.
Graphs are built by the product's own neighbour-list kernel before the timed region (the reference builds graphs
offline too) -- except in the `supercell` workload, whose step includes the graph build.

Default precision = "bf16x3": the tensor-core mode that meets the north_star tolerance (2e-3 relative) in TRAINING mode
(tests/test_gpu_bf16x3.py). The other modes are measured next to it (`by_precision`): "bf16" is ~2x faster but outside
the tolerance under batch statistics, "fp32" is the 1e-5 SIMT parity path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM_IN, DIM_RBF, NUM_LAYERS = 256, 64, 4
# algorithmic (compulsory) HBM bytes and FLOPs per edge -- per layer forward 2.10 KB / 0.524 MFLOP (K = 256
# split), per layer backward 3.16 KB / 1.05 MFLOP, edge encoder 1.04 KB fwd + 1.02 KB bwd / 0.33 + 0.66 MFLOP
ALG = {
    "train": dict(bytes_per_edge=23.1e3, flop_per_edge=7.3e6),                                   # model step fwd + bwd
    "eval": dict(bytes_per_edge=4 * 2.10e3 + 1.04e3, flop_per_edge=4 * 0.524e6 + 0.33e6),        # forward only
}
WORKLOADS = {
    "adp_train": dict(shape="adp", batch=64, mode="train", seed=2, model=dict(temperature=True, cholesky=True),
                      desc="CartNet ADP training step (fwd+bwd+Adam), batch %d crystals per GPU", config="BASELINE configs[1]"),
    "jarvis_infer": dict(shape="jarvis", batch=4096, mode="eval", seed=3, model=dict(temperature=False, cholesky=False),
                         desc="CartNet JARVIS dft_3d-shape inference (eval mode, scalar head), batch %d crystals per GPU", config="BASELINE configs[2]"),
    "mp_train": dict(shape="mp", batch=64, mode="train", seed=4, model=dict(temperature=False, cholesky=False),
                     desc="CartNet Materials-Project-shape training step (fwd+bwd+Adam, scalar head), batch %d crystals per GPU", config="BASELINE configs[3]"),
    "supercell": dict(shape="supercell", batch=1, mode="train", seed=5, model=dict(temperature=True, cholesky=True),
                      desc="CartNet large-supercell step (5000 atoms per crystal: graph build + fwd+bwd+Adam), %d crystal(s) per GPU", config="BASELINE configs[4]"),
}
PRECISIONS = ["bf16x3", "bf16", "tf32", "fp32"]
DTYPE = {"bf16x3": "bf16x3", "bf16": "bf16", "tf32": "tf32", "fp32": "f32"}

_JSON_OUT = None


def protect_stdout():
    """Rank 0 prints ONE JSON line on stdout: anything else a library writes to file descriptor 1 (NCCL's version banner
    goes there) is sent to stderr; the JSON line is written to the saved descriptor by emit()."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: str):
    out = _JSON_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops_sustained"], tensor_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src="fallback")


# ----------------------------------------------------------------------------------------- workload
def host_structures(structs, seed: int, cholesky: bool = True):
    """Collated HOST tensors of a list of synthetic crystals (cartnet_b200.synthetic.make_structures)."""
    from cartnet_b200 import synthetic
    rng = np.random.default_rng(seed + 7919)
    z = np.concatenate([s["z"] for s in structs])
    mask = z != 1
    y = synthetic.adp_targets(int(mask.sum()), rng) if cholesky else rng.standard_normal(len(structs)).astype(np.float32)
    out = dict(
        x=torch.from_numpy(z), pos=torch.from_numpy(np.concatenate([s["pos"] for s in structs])),
        cell=torch.from_numpy(np.stack([s["cell"] for s in structs])),
        natoms=torch.tensor([len(s["z"]) for s in structs], dtype=torch.int64),
        temperature=torch.tensor([s["temperature"] for s in structs], dtype=torch.float32),
        non_H_mask=torch.from_numpy(mask), y=torch.from_numpy(y))
    out["batch"] = torch.repeat_interleave(torch.arange(len(structs)), out["natoms"])
    return out


def rank_structures(shape: str, batch: int, seed: int, rank: int, world: int, device):
    """This rank's crystals of the step's GLOBAL batch (batch * world crystals from one seed). SURVEY.md 8(e): whole
    crystals per rank, balanced by EDGE count (cartnet_b200.ddp.shard_by_edges, LPT) -- the layer cost is linear in
    edges, and the step time of a data-parallel job is the slowest rank's. world == 1: all crystals, in order."""
    from cartnet_b200 import build_graph, synthetic
    from cartnet_b200.ddp import shard_by_edges
    structs = synthetic.make_structures(shape, batch * world, seed)
    if world == 1:
        return structs
    if batch == 1:
        return [structs[rank]]
    h = host_structures(structs, seed)
    gr = build_graph(h["pos"].to(device), h["cell"].to(device), h["natoms"].to(device), 5.0)
    counts = torch.bincount(h["batch"].to(device)[gr["edge_index"][1]], minlength=len(structs)).cpu().tolist()
    return [structs[i] for i in shard_by_edges(counts, world)[rank]]


def make_host_batch(structs, seed: int, device, cholesky: bool = True):
    """Synthetic crystals + graph built by the GPU neighbour-list kernel, returned as a pinned HOST batch
    (what a DataLoader with pin_memory=True hands to train.py:169)."""
    from cartnet_b200 import build_graph
    from cartnet_b200.batch import CrystalBatch
    h = host_structures(structs, seed, cholesky)
    gr = build_graph(h["pos"].to(device), h["cell"].to(device), h["natoms"].to(device), 5.0)
    h["edge_index"], h["cart_dist"], h["cart_dir"] = gr["edge_index"].cpu(), gr["cart_dist"].cpu(), gr["cart_dir"].cpu()
    # facts the data pipeline knows statically: radius_graph_pbc output is dst-sorted; the non-H atom list is fixed
    h["non_H_index"] = torch.nonzero(h["non_H_mask"]).squeeze(-1)
    b = CrystalBatch(**h).pin_memory()
    b.edges_dst_sorted = True
    return b


def shallow(b):
    from cartnet_b200.batch import CrystalBatch
    return CrystalBatch(**b.__dict__)      # forward() overwrites .x / .edge_attr on the copy only


def batch_bytes(b) -> int:
    return int(sum(v.numel() * v.element_size() for v in b.__dict__.values() if torch.is_tensor(v)))


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ----------------------------------------------------------------------------------------- reference arm
def cpu_sample_plan(wl, batch: int, override):
    """How much of the workload the CPU arm runs per step: the full batch where the oracle finishes a step in seconds
    (ADP-64: ~5 s, MP-64: <1 s), a bounded sample where it cannot (4096 JARVIS crystals, a 5000-atom supercell)."""
    if override:
        return int(override), None
    if wl["shape"] == "jarvis":
        return min(batch, 256), None
    if wl["shape"] == "supercell":
        return 1, 600          # one crystal of 600 atoms: the oracle graph build is O(n^2 C) in memory
    return batch, None


def cpu_reference_batch(wl, batch: int, seed: int, sample: int, atoms):
    """Oracle-built (CPU) batch of the first `sample` crystals of the workload's batch."""
    from cartnet_b200 import synthetic
    from oracle import fixtures
    sizes = synthetic.crystal_sizes(wl["shape"], batch, np.random.default_rng(seed))[:sample]
    if atoms:
        sizes = np.full(sample, atoms)
    return fixtures.make_oracle_batch(wl["shape"], sample, seed, sizes=sizes, cholesky=wl["model"]["cholesky"], temperature=True)


def cpu_reference_step_time(wl, hb, steps: int, warmup: int, budget_s: float = None):
    """The reference algorithm (oracle port of models/cartnet.py, eager PyTorch fp32 on CPU, all host threads):
    train = fwd + L1 + bwd + Adam, eval = forward under no_grad. Returns (s/step, graphs, edges, timed steps).
    budget_s bounds the wall time: the full ADP batch takes ~12 s per step on 16 cores, so W + K = 25 steps would run for
    five minutes; once the budget is spent (and at least one warm-up and two timed steps are done) the loop stops."""
    from oracle import cartnet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = O.OracleCartNet(DIM_IN, DIM_RBF, NUM_LAYERS, **wl["model"])
    train = wl["mode"] == "train"
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train(train)
    times = []
    t_start = time.perf_counter()
    it = -1
    while len(times) < steps:
        it += 1
        spent = time.perf_counter() - t_start
        if budget_s is not None and spent > budget_s and len(times) >= 2:
            break
        if budget_s is not None and spent > 0.4 * budget_s and it >= 1 and it < warmup:
            warmup = it                      # out of warm-up budget: start timing
        t0 = time.perf_counter()
        if train:
            opt.zero_grad(set_to_none=True)
            pred, true = model(shallow(hb))
            loss = torch.nn.functional.l1_loss(pred, true)
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                model(shallow(hb))
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times)), int(hb.natoms.numel()), int(hb.edge_index.shape[1]), len(times)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # none of the product's kernels on this arm: the graphs come from the oracle graph builder (CPU)
    sample, atoms = cpu_sample_plan(wl, args.batch, args.cpu_sample)
    hb = cpu_reference_batch(wl, args.batch, args.seed, sample, atoms)
    sec, graphs, edges, n_timed = cpu_reference_step_time(wl, hb, args.steps, args.warmup, budget_s=args.cpu_budget_s)
    cores = os.cpu_count() or 1
    val = graphs / sec
    full = graphs == args.batch and not atoms
    what = ("all %d crystals" % graphs) if full else ("%d of %d crystals%s" % (graphs, args.batch, (" of %d atoms" % atoms) if atoms else ""))
    line = {
        "impl": "reference", "metric": "graphs_per_sec", "value": val, "unit": "graphs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "edges_per_sec": edges / sec,
        "config": {"workload": (wl["desc"] % args.batch) + " -- reference algorithm (oracle port of models/cartnet.py) on the host CPU",
                   "baseline_config": wl["config"], "name": args.workload, "crystals_per_step": graphs, "edges_per_step": edges, "full_batch": full},
        "cpu_baseline": {"value": val, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "sample": "%s (%d edges) of the workload's batch per step; %d timed steps (wall budget %d s)" % (what, edges, n_timed, int(args.cpu_budget_s))},
        "e2e": {"value": val, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


# ----------------------------------------------------------------------------------------- instrumented pass
class OpTimer:
    """Times every C-ABI op group with CUDA events on the launching (current) stream during an extra,
    untimed pass; used only to attribute the step to kernels (kernel_ms_per_step, dominant_kernel)."""

    NAMES = ["edge_features", "gemm", "gemm_colstats", "gemm_tn", "colstats", "gate_center", "colsum", "edge_gate_aggregate", "node_update",
             "node_update_bwd", "edge_gate_bwd", "segment_sum", "segment_sum_pair", "dsilu_mul", "cast", "cholesky_head_fwd", "cholesky_head_bwd"]

    def __init__(self):
        from cartnet_b200 import ops
        self.ops, self.rec, self.saved = ops, [], {}

    def __enter__(self):
        for name in self.NAMES:
            fn = getattr(self.ops, name)
            self.saved[name] = fn
            setattr(self.ops, name, self._wrap(name, fn))
        return self

    def __exit__(self, *a):
        for name, fn in self.saved.items():
            setattr(self.ops, name, fn)

    def _wrap(self, name, fn):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            key, flops, nbytes = name, 0.0, 0.0
            es = lambda t: 0 if t is None else t.numel() * t.element_size()
            if name in ("gemm", "gemm_colstats"):
                A, B = a[1], a[2]
                M, K, N = A.shape[0], A.shape[1], B.shape[0]
                tag = "".join(t for t, on in (("+gather", k.get("gather0") is not None), ("+silu", k.get("act") == 1),
                                              ("+dsilu", k.get("act") == 2), ("+resid", k.get("resid") is not None),
                                              ("+colstats", name == "gemm_colstats")) if on)
                key = "gemm_nt[M=%s,N=%d,K=%d%s]" % ("E" if M > 100000 else "N", N, K, tag)
                flops = 2.0 * M * N * K
                outs = [a[4]] if name == "gemm_colstats" else [k.get(q) for q in ("z_out", "out_f32", "out_t", "resid", "z_in")]
                nbytes = es(A) + es(B) + sum(es(t) for t in outs)
            elif name == "gemm_tn":
                A, B = a[1], a[2]
                key = "gemm_tn[M=%d,N=%d,K=%s]" % (A.shape[1], B.shape[1], "E" if A.shape[0] > 100000 else "N")
                flops = 2.0 * A.shape[0] * A.shape[1] * B.shape[1]
                nbytes = es(A) + es(B)
            elif name == "edge_gate_aggregate":
                g = a[0]
                ts = a[1].element_size()
                nbytes = g.numel() * (4.0 * 2 + 2 * ts + (ts if a[13] else 0))      # e read, e' write (fp32); g, s read (T) (+T shadow of e')
            elif name == "edge_gate_bwd":
                g = a[0]
                nbytes = g.numel() * (4.0 + 7 * g.element_size())   # de read (fp32); g x2, s, dghat r+w, ds, dg (T)
            elif name in ("segment_sum", "segment_sum_pair", "colstats", "colsum"):
                nbytes = es(a[0])
            self.rec.append((key, e0, e1, flops, nbytes))
            return r
        return inner

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for key, e0, e1, flops, nbytes in self.rec:
            d = agg.setdefault(key, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
            d["ms"] += e0.elapsed_time(e1)
            d["n"] += 1
            d["flops"] += flops
            d["bytes"] += nbytes
        return agg


# ----------------------------------------------------------------------------------------- main arm
def run_ours(args, wl):
    import torch.distributed as dist

    import cartnet_b200
    from cartnet_b200 import build_graph, ops
    from cartnet_b200 import cartnet as CN
    from cartnet_b200.ddp import FlatGradAllReduce, broadcast_module

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    train = wl["mode"] == "train"
    chol = wl["model"]["cholesky"]
    with_graph = wl["shape"] == "supercell"       # this workload's step includes the neighbour-list build

    nb = 2
    # step i of every rank works on its shard of global batch i (args.batch * world crystals, seed + i)
    host_batches = [make_host_batch(rank_structures(wl["shape"], args.batch, args.seed + i, rank, world, dev), args.seed + 1000 * rank + i, dev, chol)
                    for i in range(nb)]
    dev_batches = [shallow(hb).to(dev) for hb in [b.clone() for b in host_batches]]
    graphs_step = args.batch                     # per rank on average: the global batch has args.batch * world crystals
    edges_step = float(np.mean([b.num_edges for b in host_batches]))
    nodes_step = float(np.mean([b.num_nodes for b in host_batches]))

    def build_model(precision):
        torch.manual_seed(0)
        m = cartnet_b200.CartNet(DIM_IN, DIM_RBF, NUM_LAYERS, precision=precision, **wl["model"]).to(dev)
        broadcast_module(m, 0)
        m.train(train)
        return m

    model = build_model(args.precision)
    sync = FlatGradAllReduce(model.parameters(), direct=True) if train else None
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True) if train else None      # one multi-tensor launch
    for b in dev_batches:
        CN.get_plan(b)                      # graph plans are per-batch preprocessing (cached by edge_index identity)

    graph_ms = None
    if with_graph:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b = dev_batches[0]
        build_graph(b.pos, b.cell, b.natoms, 5.0)
        e0.record()
        for _ in range(5):
            build_graph(b.pos, b.cell, b.natoms, 5.0)
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / 5

    def regraph(b):
        """supercell workload: positions -> periodic radius graph on the device, every step"""
        gr = build_graph(b.pos, b.cell, b.natoms, 5.0)
        b.edge_index, b.cart_dist, b.cart_dir = gr["edge_index"], gr["cart_dist"], gr["cart_dir"]
        return b

    def make_step(model, sync, opt):
        def step(b, collective=True):
            if with_graph:
                b = regraph(b)
            if not train:
                with torch.no_grad():
                    pred, _ = model(b)
                return pred
            sync.zero()
            pred, true = model(b)
            loss = cartnet_b200.compute_loss(pred, true)[0]      # MAE of (MAE, MSE), train/metrics.py:15-28, cfg.loss = "MAE"
            loss.backward()
            if collective:
                sync.allreduce_mean()
            opt.step()
            return loss
        return step

    step = make_step(model, sync, opt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident inputs ("value")
    for i in range(args.warmup):
        step(shallow(dev_batches[i % nb]))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    ms = timed(lambda i: step(shallow(dev_batches[i % nb])), args.steps)
    launches = ops.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms / args.steps
    value = world * graphs_step / (ms_step * 1e-3)

    if os.environ.get("CARTNET_BENCH_PROFILE_ONLY") == "1":      # ncu launch lists (scripts/step_traffic.py): the timed loop is all that is wanted
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_per_step": ms_step, "note": "not a bench line"}))
        return

    # ---- end to end through the public API: pinned host batch -> .to(device) -> model -> result read back on the host
    h2d = int(np.mean([batch_bytes(b) for b in host_batches]))

    from cartnet_b200 import DevicePrefetcher
    n_e2e_warm = max(4, args.warmup)         # the copy stream's allocator pool reaches its steady state within a few batches
    feed = DevicePrefetcher((host_batches[i % nb] for i in range(n_e2e_warm + args.steps)), dev, eager=False)

    def e2e_step(i):
        b = next(feed)                       # pinned host batch -> device (copy + graph plan issued one step ahead)
        out = step(b)
        feed.prefetch_next()                 # the next batch's copies are issued behind this step's launches, not in front of them
        if train:
            return float(out.item())         # device -> host read of the step's result, every step
        return out.float().cpu()             # inference: the predictions themselves

    for i in range(n_e2e_warm):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * graphs_step / (ms_e2e * 1e-3)
    d2h = 4
    if not train:
        d2h = int(step(shallow(dev_batches[0])).numel() * 4)

    # ---- the same loop fed from a DEVICE-RESIDENT data set (cartnet_b200.DeviceDataset: batches assembled on the GPU by one
    # launch); per step only the crystal-id list crosses PCIe. Reported next to `e2e`, not instead of it.
    e2e_dev = None
    if world == 1 and not with_graph:
        from cartnet_b200 import DeviceDataset
        dsets = [DeviceDataset.from_batch(hb, dev) for hb in host_batches]
        ids = list(range(args.batch))

        def dev_step(i):
            out = step(dsets[i % nb].collate(ids))
            return float(out.item()) if train else out.float().cpu()

        for i in range(3):
            dev_step(i)
        ms_dev = timed(dev_step, args.steps) / args.steps
        e2e_dev = {"value": graphs_step / (ms_dev * 1e-3), "unit": "graphs/s", "ms_per_step": ms_dev,
                   "h2d_bytes_per_step": 4 * (args.batch + 4 * (args.batch + 1)), "d2h_bytes_per_step": d2h,
                   "note": "data set resident in HBM, batch assembled on the device (cartnet_collate), result read back every step"}
        del dsets

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- the other precision modes on the same batches (short loops; N = 1 only)
    by_precision = {args.precision: {"ms_per_step": ms_step, "graphs_per_sec": value / world, "steps": args.steps}}
    if world == 1 and not args.no_by_precision:
        for prec in PRECISIONS:
            if prec == args.precision:
                continue
            m2 = build_model(prec)
            s2 = FlatGradAllReduce(m2.parameters(), direct=True) if train else None
            o2 = torch.optim.Adam(m2.parameters(), lr=1e-3, fused=True) if train else None
            st2 = make_step(m2, s2, o2)
            k = 3 if prec == "fp32" else 5
            for i in range(2):
                st2(shallow(dev_batches[i % nb]))
            t = timed(lambda i: st2(shallow(dev_batches[i % nb])), k) / k
            by_precision[prec] = {"ms_per_step": t, "graphs_per_sec": graphs_step / (t * 1e-3), "steps": k}
            del m2, s2, o2, st2
        notes = {"bf16x3": "hi|lo bf16 pairs, 3 tcgen05 MMAs per product: within 2e-3 in training AND eval mode (default)",
                 "bf16": "plain bf16 operands: within 2e-3 in eval mode on the ADP case only; training mode 3e-2 (edge BatchNorm amplifies operand rounding)",
                 "tf32": "tf32 operands: within 2e-3 in eval mode on every golden case; training mode 5e-3",
                 "fp32": "fp32 SIMT parity path: within 1e-5"}
        for p in by_precision:
            by_precision[p]["tolerance"] = notes[p]

    # ---- kernel attribution (extra instrumented pass, not part of any reported time)
    pk = peaks()
    from cartnet_b200 import functional as CF
    CF.USE_NATIVE_LAYER = False              # same kernels, issued one by one from Python so that each can be timed
    step(shallow(dev_batches[0]), collective=False)                  # warm pass: this mode's allocation pattern (no cudaMalloc inside a timed bracket)
    torch.cuda.synchronize()
    with OpTimer() as ot:
        for i in range(2):
            step(shallow(dev_batches[i % nb]), collective=False)     # rank 0 only: no collective in this pass
        agg = ot.summary()
    CF.USE_NATIVE_LAYER = True
    tot_ms = sum(d["ms"] for d in agg.values())
    top_key, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    breakdown = {k: round(v["ms"] / 2, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- roofline: SURVEY 8(d) algorithmic bytes of one step / measured step time, against the measured copy peak
    alg = ALG[wl["mode"]]
    step_s = ms_step * 1e-3
    alg_bytes = alg["bytes_per_edge"] * edges_step
    achieved = alg_bytes / step_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.isfile(tp):
        traffic = json.load(open(tp)).get("%s/%s/batch%d" % (args.workload, args.precision, args.batch))   # DRAM bytes per step from the committed ncu window
    roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"], "traffic": traffic,
            "definition": "SURVEY 8(d) algorithmic bytes per edge x edges per step / ms_per_step (whole model step: edge encoder + 4 fused-layer units%s)" % (" fwd+bwd" if train else ", forward"),
            "algorithmic_bytes_per_edge": alg["bytes_per_edge"], "algorithmic_bytes_per_step": alg_bytes,
            "tensor_frac": alg["flop_per_edge"] * edges_step / step_s / (pk["tensor"] * 1e12), "algorithmic_flop_per_edge": alg["flop_per_edge"],
            "peak_source": pk["src"], "edges_per_sec_per_gpu": edges_step / step_s,
            "dominant_kernel": {"kernel": top_key, "launches_per_step": top["n"] / 2, "avg_launch_ms": top["ms"] / top["n"],
                                "share_of_step": top["ms"] / tot_ms,
                                "kernel_bw_util": (top["bytes"] / (top["ms"] * 1e-3) / 1e9 / pk["hbm"]) if top["bytes"] else None,
                                "kernel_bytes_moved_per_launch": top["bytes"] / top["n"],
                                "kernel_tensor_util": top["flops"] / (top["ms"] * 1e-3) / 1e12 / pk["tensor"]}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample, atoms = cpu_sample_plan(wl, args.batch, args.cpu_sample)
        hb = cpu_reference_batch(wl, args.batch, args.seed, sample, atoms)
        sec, g, e, _ = cpu_reference_step_time(wl, hb, 2, 1)
        cpu = {"value": g / sec, "unit": "graphs/s", "cores": os.cpu_count() or 1, "kind": "port", "edges_per_sec": e / sec,
               "sample": "%d crystals (%d edges)%s of the workload's batch, %s, 1 warm-up + mean of 2" % (
                   g, e, " = the full batch" if (g == args.batch and not atoms) else "", "fwd+bwd+Adam" if train else "eval forward")}

    act_gb = edges_step * (DIM_IN * 4 * 8) * NUM_LAYERS / 1e9
    line = {
        "metric": "graphs_per_sec", "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE[args.precision], "data": "synthetic",
        "edges_per_sec": world * edges_step / step_s,
        "config": {"workload": wl["desc"] % args.batch, "baseline_config": wl["config"], "name": args.workload,
                   "crystals_per_gpu": args.batch, "atoms_per_gpu": nodes_step, "edges_per_gpu": edges_step, "radius": 5.0,
                   "dim_in": DIM_IN, "dim_rbf": DIM_RBF, "num_layers": NUM_LAYERS, "precision": args.precision,
                   "parallelism": (("global batch of %d crystals sharded as whole crystals per GPU, balanced by edge count (LPT); one NCCL all-reduce of the flat gradient per step" % (args.batch * world)) if train else "independent replicas, no collective") if world > 1 else "single GPU",
                   "l2": "per-step working set ~%.1f GB of activations >> 126 MB L2; %d distinct batches cycled" % (act_gb, nb)},
        "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                "result_read": "blocking loss.item() every step" if train else "predictions copied to the host every step"},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
        "clocks": clocks, "roofline": roof, "by_precision": by_precision, "kernel_ms_per_step": breakdown,
    }
    if graph_ms is not None:
        line["graph_build_ms"] = graph_ms
    if e2e_dev is not None:
        line["e2e_device_dataset"] = e2e_dev
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="adp_train", choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("CARTNET_BENCH_PRECISION", "bf16x3"), choices=PRECISIONS)
    ap.add_argument("--batch", type=int, default=None, help="crystals per GPU (default: the workload's)")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=None, help="crystals per step for the CPU reference (default: the full batch where the oracle finishes in seconds)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="--impl reference: wall-time bound of the CPU loop (>= 1 warm-up + 2 timed steps always run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-by-precision", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.batch = args.batch or wl["batch"]
    args.seed = wl["seed"] if args.seed is None else args.seed
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    protect_stdout()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
        run_ours(args, wl)


if __name__ == "__main__":
    main()
