#!/usr/bin/env python
"""Times the tcgen05 NT / TN GEMMs in isolation at the ADP-64 shapes (CUDA events, inputs > L2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cartnet_b200 import ops  # noqa: E402

E, N = 687000, 12400
prec = {"bf16": ops.PREC_BF16, "tf32": ops.PREC_TF32, "fp32": ops.PREC_FP32, "bf16x3": ops.PREC_BF16X3}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = ops.t_dtype(prec)
dev = "cuda"


def tt(x):
    """fp32 values -> T-typed operand buffer of the mode"""
    return ops.cast(x.contiguous(), prec) if prec in (ops.PREC_TF32, ops.PREC_BF16X3) else x.to(T)


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def nt(M, Nn, K, mode):
    A = tt(torch.randn(M, K, device=dev))
    B = tt(torch.randn(Nn, K, device=dev) * K ** -0.5)
    kw = {}
    if mode == "f32out":
        kw = dict(bias=torch.randn(Nn, device=dev), out_f32=torch.empty(M, Nn, device=dev))
    elif mode == "tout":
        kw = dict(out_t=torch.empty(M, Nn, device=dev, dtype=T))
    elif mode.startswith("gather_silu"):
        P = torch.randn(N, 2 * Nn, device=dev)
        P = P if prec == ops.PREC_BF16X3 else tt(P)          # bf16x3: gathered operands are plain fp32
        dst = torch.sort(torch.randint(0, N, (M,), device=dev))[0].to(torch.int32)
        src = torch.randint(0, N, (M,), device=dev, dtype=torch.int32)
        if mode == "gather_silu_same":      # both gathers hit one row: L1-resident (isolates the gather path)
            dst, src = torch.zeros_like(dst), torch.zeros_like(src)
        elif mode == "gather_silu_local":   # src close to dst (sorted): L1-friendly
            src = dst.clone()
        kw = dict(bias=torch.randn(Nn, device=dev), gather0=P[:, :Nn], gidx0=dst, gather1=P[:, Nn:], gidx1=src,
                  z_out=torch.empty(M, Nn, device=dev, dtype=ops.z_dtype(prec)), act=ops.ACT_SILU, out_t=torch.empty(M, Nn, device=dev, dtype=T))
    elif mode == "resid":
        kw = dict(resid=torch.randn(M, Nn, device=dev), out_f32=torch.empty(M, Nn, device=dev))
    elif mode == "bias_tout":
        kw = dict(bias=torch.randn(Nn, device=dev), out_t=torch.empty(M, Nn, device=dev, dtype=T))
    elif mode == "dsilu":
        kw = dict(act=ops.ACT_MUL_DSILU, z_in=(torch.randn(M, Nn, device=dev).half() if prec == ops.PREC_BF16X3 else tt(torch.randn(M, Nn, device=dev))), out_t=torch.empty(M, Nn, device=dev, dtype=T))
    ms = timeit(lambda: ops.gemm(prec, A, B, **kw), reps)
    print("NT  M=%7d N=%4d K=%4d %-12s %8.3f ms  %7.1f TFLOP/s" % (M, Nn, K, mode, ms, 2.0 * M * Nn * K / ms / 1e9))


def tn(K, M, Nn):
    A = tt(torch.randn(K, M, device=dev))
    B = tt(torch.randn(K, Nn, device=dev))
    ms = timeit(lambda: ops.gemm_tn(prec, A, B), reps)
    print("TN  K=%7d M=%4d N=%4d              %8.3f ms  %7.1f TFLOP/s" % (K, M, Nn, ms, 2.0 * M * Nn * K / ms / 1e9))


which = sys.argv[3] if len(sys.argv) > 3 else "all"
if which in ("all", "nt"):
    nt(E, 256, 256, "tout")
    nt(E, 256, 256, "f32out")
    nt(E, 256, 256, "dsilu")
    nt(E, 512, 256, "gather_silu")
    nt(E, 256, 512, "f32out")
    nt(E, 256, 512, "resid")
    nt(E, 256, 256, "bias_tout")
    nt(E // 8, 256, 256, "tout")
if which in ("all", "tn"):
    tn(E, 256, 256)
    tn(E, 512, 256)
if which == "one":
    nt(E, 256, 256, "tout")
if which == "hot":      # the two kernels furthest from their bound (ncu targets)
    nt(E, 512, 256, "gather_silu")
    nt(E, 256, 512, "resid")
if which == "gs":
    nt(E, 512, 256, "gather_silu")
    nt(E, 512, 256, "gather_silu_local")
    nt(E, 512, 256, "gather_silu_same")
