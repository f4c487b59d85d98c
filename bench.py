#!/usr/bin/env python
"""bench.py -- CartNet ADP training-step throughput (BASELINE.json: crystal graphs/s & edges/s, fwd+bwd,
ADP shape) on N B200s of one node, with the reference CPU path timed beside it.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1      # the reference algorithm on the host cores

Workload (config.workload): BASELINE configs[1] -- "CartNet ADP training step, batch 64 crystals, 1xB200":
64 synthetic ADP-shaped crystals per GPU (lognormal sizes, mean ~194 atoms, 9.5 A^3/atom, radius 5 A),
random-init CartNet(256, 64, 4 layers, Cholesky head), one step = forward + L1 loss + backward +
gradient all-reduce (N > 1) + Adam. Weak scaling: the global batch is 64 x N crystals, sharded as whole crystals per rank balanced by edge count.
Graphs are built by the product's own neighbour-list kernel before the timed region (the reference builds
graphs offline too).
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
KB_PER_EDGE_STEP = 23.1e3        # algorithmic HBM bytes per edge per model step (fwd+bwd)
FLOP_PER_EDGE_STEP = 7.3e6       # algorithmic FLOPs per edge per step with the K=256 split


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
def host_structures(structs, seed: int):
    """Collated HOST tensors of a list of synthetic crystals (cartnet_b200.synthetic.make_structures)."""
    from cartnet_b200 import synthetic
    rng = np.random.default_rng(seed + 7919)
    z = np.concatenate([s["z"] for s in structs])
    mask = z != 1
    out = dict(
        x=torch.from_numpy(z), pos=torch.from_numpy(np.concatenate([s["pos"] for s in structs])),
        cell=torch.from_numpy(np.stack([s["cell"] for s in structs])),
        natoms=torch.tensor([len(s["z"]) for s in structs], dtype=torch.int64),
        temperature=torch.tensor([s["temperature"] for s in structs], dtype=torch.float32),
        non_H_mask=torch.from_numpy(mask), y=torch.from_numpy(synthetic.adp_targets(int(mask.sum()), rng)))
    out["batch"] = torch.repeat_interleave(torch.arange(len(structs)), out["natoms"])
    return out


def rank_structures(batch: int, seed: int, rank: int, world: int, device):
    """This rank's crystals of the step's GLOBAL batch (batch * world crystals from one seed). SURVEY.md 8(e): whole
    crystals per rank, balanced by EDGE count (cartnet_b200.ddp.shard_by_edges, LPT) -- the layer cost is linear in
    edges, and the step time of a data-parallel job is the slowest rank's. world == 1: all crystals, in order."""
    from cartnet_b200 import build_graph, synthetic
    from cartnet_b200.ddp import shard_by_edges
    structs = synthetic.make_structures("adp", batch * world, seed)
    if world == 1:
        return structs
    h = host_structures(structs, seed)
    gr = build_graph(h["pos"].to(device), h["cell"].to(device), h["natoms"].to(device), 5.0)
    counts = torch.bincount(h["batch"].to(device)[gr["edge_index"][1]], minlength=len(structs)).cpu().tolist()
    return [structs[i] for i in shard_by_edges(counts, world)[rank]]


def make_host_batch(structs, seed: int, device):
    """Synthetic crystals + graph built by the GPU neighbour-list kernel, returned as a pinned HOST batch
    (what a DataLoader with pin_memory=True hands to train.py:169)."""
    from cartnet_b200 import build_graph
    from cartnet_b200.batch import CrystalBatch
    h = host_structures(structs, seed)
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
            self._stop.wait(0.2)

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
def cpu_reference_step_time(hb, sample_crystals: int, steps: int, warmup: int):
    """The reference algorithm (oracle port of models/cartnet.py, eager PyTorch fp32 on CPU, all host threads)
    on the first `sample_crystals` crystals of the batch: fwd + L1 + bwd + Adam. Returns (s/step, graphs, edges)."""
    from cartnet_b200.batch import CrystalBatch
    from oracle import cartnet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    nat = hb.natoms[:sample_crystals]
    n = int(nat.sum())
    emask = hb.edge_index[1] < n
    sub = CrystalBatch(x=hb.x[:n].clone(), batch=hb.batch[:n].clone(), natoms=nat.clone(), temperature=hb.temperature[:sample_crystals].clone(),
                       non_H_mask=hb.non_H_mask[:n].clone(), y=hb.y[: int(hb.non_H_mask[:n].sum())].clone(),
                       edge_index=hb.edge_index[:, emask].clone(), cart_dist=hb.cart_dist[emask].clone(), cart_dir=hb.cart_dir[emask].clone())
    torch.manual_seed(0)
    model = O.OracleCartNet(DIM_IN, DIM_RBF, NUM_LAYERS)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model.train()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        pred, true = model(shallow(sub))
        loss = torch.nn.functional.l1_loss(pred, true)
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times)), sample_crystals, int(emask.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # none of the product's kernels on this arm: the sample's graphs come from the oracle graph builder (CPU)
    from cartnet_b200 import synthetic
    from oracle import fixtures
    sizes = synthetic.crystal_sizes("adp", args.batch, np.random.default_rng(args.seed))[:args.cpu_sample]
    hb = fixtures.make_oracle_batch("adp", args.cpu_sample, args.seed, sizes=sizes)
    sec, graphs, edges = cpu_reference_step_time(hb, args.cpu_sample, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    val = graphs / sec
    line = {
        "impl": "reference", "metric": "graphs_per_sec", "value": val, "unit": "graphs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "edges_per_sec": edges / sec,
        "config": {"workload": "CartNet ADP training step (fwd+bwd+Adam), ADP-shaped synthetic crystals, reference algorithm on host CPU",
                   "sample_crystals": graphs, "sample_edges": edges},
        "cpu_baseline": {"value": val, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "sample": "%d crystals (%d edges) of the ADP-64 batch per step" % (graphs, edges)},
        "e2e": {"value": val, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


# ----------------------------------------------------------------------------------------- instrumented pass
class OpTimer:
    """Times every C-ABI op group with CUDA events on the launching (current) stream during an extra,
    untimed pass; used only to attribute the step to kernels for the roofline line."""

    NAMES = ["edge_features", "gemm", "gemm_tn", "colstats", "colsum", "edge_gate_aggregate", "node_update",
             "node_update_bwd", "edge_gate_bwd", "segment_sum", "dsilu_mul", "cast"]

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
            if name == "gemm":
                A, B = a[1], a[2]
                M, K, N = A.shape[0], A.shape[1], B.shape[0]
                tag = "".join(t for t, on in (("+gather", k.get("gather0") is not None), ("+silu", k.get("act") == 1),
                                              ("+dsilu", k.get("act") == 2), ("+resid", k.get("resid") is not None)) if on)
                key = "gemm_nt[M=%s,N=%d,K=%d%s]" % ("E" if M > 100000 else "N", N, K, tag)
                flops = 2.0 * M * N * K
                nbytes = es(A) + es(B) + sum(es(k.get(q)) for q in ("z_out", "out_f32", "out_t", "resid", "z_in"))
            elif name == "gemm_tn":
                A, B = a[1], a[2]
                key = "gemm_tn[M=%d,N=%d,K=%s]" % (A.shape[1], B.shape[1], "E" if A.shape[0] > 100000 else "N")
                flops = 2.0 * A.shape[0] * A.shape[1] * B.shape[1]
                nbytes = es(A) + es(B)
            elif name == "edge_gate_aggregate":
                g = a[0]
                ts = a[1].element_size()
                nbytes = g.numel() * (4.0 * 3 + 2 * ts + (ts if a[12] != 0 else 0))      # g,e read, e' write (fp32); s read, gn write (T) (+T shadow)
            elif name == "edge_gate_bwd":
                g = a[0]
                nbytes = g.numel() * (4.0 + 7 * g.element_size())   # de read (fp32); gn x2, s, dghat r+w, ds, dg (T)
            elif name == "segment_sum":
                nbytes = es(a[0])
            elif name in ("colstats", "colsum"):
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
def run_ours(args):
    import torch.distributed as dist

    import cartnet_b200
    from cartnet_b200 import ops
    from cartnet_b200.ddp import FlatGradAllReduce, broadcast_module

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus

    nb = 2
    # step i of every rank works on its shard of global batch i (args.batch * world crystals, seed + i)
    host_batches = [make_host_batch(rank_structures(args.batch, args.seed + i, rank, world, dev), args.seed + 1000 * rank + i, dev)
                    for i in range(nb)]
    dev_batches = [shallow(hb).to(dev) for hb in [b.clone() for b in host_batches]]
    graphs_step = args.batch                     # per rank on average: the global batch has args.batch * world crystals
    edges_step = float(np.mean([b.num_edges for b in host_batches]))
    nodes_step = float(np.mean([b.num_nodes for b in host_batches]))

    torch.manual_seed(0)
    model = cartnet_b200.CartNet(DIM_IN, DIM_RBF, NUM_LAYERS, precision=args.precision).to(dev)
    broadcast_module(model, 0)
    sync = FlatGradAllReduce(model.parameters(), direct=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)      # one multi-tensor launch
    model.train()
    from cartnet_b200 import cartnet as CN
    for b in dev_batches:
        CN.get_plan(b)                      # graph plans are per-batch preprocessing (cached by edge_index identity)

    def step(b, collective=True):
        sync.zero()
        pred, true = model(b)
        loss = torch.nn.functional.l1_loss(pred, true)
        loss.backward()
        if collective:
            sync.allreduce_mean()
        opt.step()
        return loss

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

    # ---- end to end through the public API: pinned host batch -> .to(device) -> model -> loss.item()
    h2d = int(np.mean([batch_bytes(b) for b in host_batches]))

    from cartnet_b200 import DevicePrefetcher
    n_e2e_warm = max(4, args.warmup)         # the copy stream's allocator pool reaches its steady state within a few batches
    feed = DevicePrefetcher((host_batches[i % nb] for i in range(n_e2e_warm + args.steps)), dev)

    def e2e_step(i):
        b = next(feed)                       # pinned host batch -> device (copy + graph plan issued one step ahead)
        loss = step(b)
        return float(loss.item())            # device -> host read of the step's result, every step

    for i in range(n_e2e_warm):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * graphs_step / (ms_e2e * 1e-3)

    # same loop with the loss read back asynchronously (cartnet_b200.DeferredScalars: D2H copy into pinned memory every
    # step, consumed one step later, the last one inside the timed region) -- reported next to the blocking variant above
    from cartnet_b200 import DeferredScalars
    feed = DevicePrefetcher((host_batches[i % nb] for i in range(n_e2e_warm + args.steps)), dev)
    losses = DeferredScalars(depth=2)

    def e2e_async_step(i, last=args.steps - 1):
        losses.push(step(next(feed)))
        if i == last:
            losses.drain()

    for i in range(n_e2e_warm):              # the host now runs one step ahead: let the caching allocator grow to that
        e2e_async_step(i, last=n_e2e_warm - 1)
    ms_e2e_async = timed(e2e_async_step, args.steps) / args.steps

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline attribution (extra instrumented pass, not part of any reported time)
    pk = peaks()
    from cartnet_b200 import functional as CF
    CF.USE_NATIVE_LAYER = False              # same kernels, issued one by one from Python so that each can be timed
    with OpTimer() as ot:
        for i in range(2):
            step(shallow(dev_batches[i % nb]), collective=False)     # rank 0 only: no collective in this pass
        agg = ot.summary()
    tot_ms = sum(d["ms"] for d in agg.values())
    top_key, top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    # which roof binds the dominant kernel: its arithmetic intensity (algorithmic FLOPs / algorithmic HBM bytes per launch)
    # against the ridge of the measured peaks; the other fraction is reported next to it
    t_frac = top["flops"] / (top["ms"] * 1e-3) / 1e12 / pk["tensor"]
    h_frac = top["bytes"] / (top["ms"] * 1e-3) / 1e9 / pk["hbm"]
    ridge = pk["tensor"] * 1e12 / (pk["hbm"] * 1e9)
    if top["flops"] > 0 and (top["bytes"] <= 0 or top["flops"] / top["bytes"] >= ridge):
        roof = {"bound": "tensor", "achieved": t_frac * pk["tensor"], "peak": pk["tensor"], "unit": "TFLOP/s", "frac": t_frac}
    else:
        roof = {"bound": "hbm", "achieved": h_frac * pk["hbm"], "peak": pk["hbm"], "unit": "GB/s", "frac": h_frac}
    roof.update({"flop_per_byte": (top["flops"] / top["bytes"]) if top["bytes"] > 0 else None, "ridge_flop_per_byte": ridge,
                 "tensor_frac": t_frac, "hbm_frac": h_frac, "algorithmic_bytes_per_launch": top["bytes"] / top["n"]})
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.isfile(tp) and args.precision == "bf16" and args.batch == 64:
        traffic = json.load(open(tp)).get(top_key)      # DRAM bytes per launch of this op from the committed ncu capture
    roof.update({"traffic": traffic, "kernel": top_key, "launches_per_step": top["n"] / 2, "avg_launch_ms": top["ms"] / top["n"],
                 "share_of_step": top["ms"] / tot_ms, "peak_source": pk["src"] + (" (sustained bf16)" if roof["bound"] == "tensor" else "")})
    breakdown = {k: round(v["ms"] / 2, 4) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    step_s = ms_step * 1e-3
    step_roof = {
        "edges_per_sec_per_gpu": edges_step / step_s,
        "hbm_frac_model_step": KB_PER_EDGE_STEP * edges_step / step_s / (pk["hbm"] * 1e9),
        "tensor_frac_model_step": FLOP_PER_EDGE_STEP * edges_step / step_s / (pk["tensor"] * 1e12),
        "algorithmic_bytes_per_edge": KB_PER_EDGE_STEP, "algorithmic_flop_per_edge": FLOP_PER_EDGE_STEP}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, g, e = cpu_reference_step_time(host_batches[0], args.cpu_sample, 2, 1)
        cpu = {"value": g / sec, "unit": "graphs/s", "cores": os.cpu_count() or 1, "kind": "port", "edges_per_sec": e / sec,
               "sample": "%d crystals (%d edges) of the ADP-64 batch, fwd+bwd+Adam, 1 warm-up + mean of 2" % (g, e)}

    act_gb = edges_step * (DIM_IN * 4 * 8) * NUM_LAYERS / 1e9
    line = {
        "metric": "graphs_per_sec", "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3", "bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.precision], "data": "synthetic",
        "edges_per_sec": world * edges_step / step_s,
        "config": {"workload": "CartNet ADP training step (fwd+bwd+Adam), batch %d crystals per GPU" % args.batch,
                   "crystals_per_gpu": args.batch, "atoms_per_gpu": nodes_step, "edges_per_gpu": edges_step, "radius": 5.0,
                   "dim_in": DIM_IN, "dim_rbf": DIM_RBF, "num_layers": NUM_LAYERS, "precision": args.precision,
                   "parallelism": "global batch of %d crystals sharded as whole crystals per GPU, balanced by edge count (LPT); one NCCL all-reduce of the flat gradient per step" % (args.batch * world) if world > 1 else "single GPU",
                   "l2": "per-step working set ~%.1f GB of activations >> 126 MB L2; %d distinct batches cycled" % (act_gb, nb)},
        "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e,
                "loss_read": "blocking loss.item() every step"},
        "e2e_async_loss": {"value": world * graphs_step / (ms_e2e_async * 1e-3), "unit": "graphs/s", "ms_per_step": ms_e2e_async,
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                           "loss_read": "async D2H into pinned memory every step, consumed one step later (DeferredScalars)"},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
        "clocks": clocks, "roofline": roof, "step_roofline": step_roof, "kernel_ms_per_step": breakdown,
    }
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
    ap.add_argument("--precision", default=os.environ.get("CARTNET_BENCH_PRECISION", "bf16"), choices=["bf16x3", "bf16", "tf32", "fp32"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=4, help="crystals per step for the CPU reference (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()
