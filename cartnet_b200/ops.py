"""Tensor-level wrappers over the C ABI (one function per entry point group).

Each wrapper checks devices/dtypes/strides, allocates outputs with torch (the library never
allocates), passes raw device pointers plus the current CUDA stream, and raises on a non-zero
return code. There is no CPU path here: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_MUL_DSILU, ACT_NONE, ACT_SILU, PREC_BF16, PREC_BF16X3, PREC_FP32, PREC_TF32  # noqa: F401

EPS_BN = 1e-5


def t_dtype(prec: int) -> torch.dtype:
    """torch dtype of the buffers that hold GEMM operands: bf16 for PREC_BF16; 4-byte words for PREC_FP32 (SIMT),
    PREC_TF32 (tcgen05 kind::tf32) and PREC_BF16X3. In PREC_BF16X3 the words are OPAQUE: each run of 64 elements
    (256 bytes) holds 64 bf16 high parts followed by 64 bf16 low parts (value = hi + lo, ~16 mantissa bits), the
    layout the split-precision tcgen05 GEMMs read with TMA (csrc/common.cuh::bf16p_t). Only the library's kernels
    may interpret such a buffer; column slices must start at multiples of 64."""
    return torch.bfloat16 if prec == PREC_BF16 else torch.float32


def z_dtype(prec: int) -> torch.dtype:
    """torch dtype of stored pre-activations (`z_out` / `z_in` of `gemm`, `z` of `dsilu_mul`): the operand dtype, except
    fp16 in PREC_BF16X3 -- a pre-activation is only re-read to evaluate silu'(z), which 2^-11 relative rounding moves by
    at most 1.2e-4 (csrc/common.cuh::ZOf); stores saturate."""
    return torch.float16 if prec == PREC_BF16X3 else t_dtype(prec)


def f32_storage(prec: int) -> bool:
    """dtype of T-typed buffers is torch.float32 (fp32 and tf32 modes)"""
    return prec != PREC_BF16


def needs_shadow(prec: int) -> bool:
    """the tensor-core modes keep separate T-typed operand copies of x / e (bf16, or fp32 words rounded to tf32);
    in fp32 mode the fp32 tensors themselves are the operands"""
    return prec != PREC_FP32


def launch_count() -> int:
    return int(_lib.load().cartnet_launch_count())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str, rowmajor: bool = True):
    if not t.is_cuda:
        raise RuntimeError("cartnet_b200: %s must be a CUDA tensor (no CPU fallback exists)" % name)
    if t.dtype != dtype:
        raise TypeError("cartnet_b200: %s must be %s, got %s" % (name, dtype, t.dtype))
    if rowmajor and t.dim() >= 1 and t.numel() > 0 and t.stride(-1) != 1:
        raise ValueError("cartnet_b200: %s must have unit stride in its last dimension" % name)
    return t


def _ld2(t: torch.Tensor) -> int:
    if t.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if t.shape[0] > 1:
        return int(t.stride(0))
    return int(max(t.shape[1], t.stride(0)))


_scratch = {}


def _partial(device, nbytes: int) -> torch.Tensor:
    """fp64 scratch for the fixed-order column reductions (stream-ordered reuse: one buffer per device AND stream, so
    that work issued on a side stream never shares scratch with the main stream)."""
    key = (device, _stream(), "partial")
    buf = _scratch.get(key)
    if buf is None or buf.numel() * 8 < nbytes:
        buf = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=device)
        _scratch[key] = buf
    return buf


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device, _stream(), "ws")
    buf = _scratch.get(key)
    if buf is None or buf.numel() * 4 < nbytes:
        buf = torch.empty((nbytes + 3) // 4 + 64, dtype=torch.float32, device=device)
        _scratch[key] = buf
    return buf


# ----------------------------------------------------------------------------- graph build
def nlist_build(pos, cell, natoms, radius: float, pbc_mask: int = 7, batch_max_reps: bool = True,
                want_cart: bool = True, want_i32: bool = False, cells: bool = True):
    """Periodic radius graph (dataset/utils.py:57-237). Returns dict with edge_index [2,E] i64,
    unit_cell, dist, direction (+ cart_dist, cart_dir, src32, dst32, row_ptr when requested).
    cells=True: the cell-list kernels (large crystals are binned, small ones keep the all-pairs scan inside the same
    launches); cells=False: the all-pairs kernels for every crystal. Both give the identical, bit-exact result."""
    lib = _lib.load()
    pos = _req(pos.contiguous(), torch.float32, "pos")
    cell = _req(cell.reshape(-1, 3, 3).contiguous(), torch.float32, "cell")
    dev = pos.device
    natoms = natoms.to(device=dev, dtype=torch.int64).reshape(-1)
    B, N = int(natoms.numel()), int(pos.shape[0])
    crystal_ptr = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    crystal_ptr[1:] = torch.cumsum(natoms, 0).to(torch.int32)
    node_crystal = torch.repeat_interleave(torch.arange(B, device=dev, dtype=torch.int32), natoms, output_size=N)
    reps = torch.empty(B, 3, dtype=torch.int32, device=dev)
    reps_max = torch.zeros(3, dtype=torch.int32, device=dev)
    st = _stream()
    _lib.check(lib.cartnet_nlist_reps(_p(cell), B, float(radius), int(pbc_mask), _p(reps), _p(reps_max), st), "nlist_reps")
    reps_arg, stride = (reps_max, 0) if batch_max_reps else (reps, 3)
    r2 = C.c_float(float(radius) * float(radius)).value      # double product rounded to fp32 (utils.py:202)
    row_count = torch.empty(N, dtype=torch.int32, device=dev)
    ws = None
    if cells and N > 0:
        ws = torch.empty(int(lib.cartnet_nlist_cells_workspace(N, B)) // 4, dtype=torch.int32, device=dev)
        _lib.check(lib.cartnet_nlist_cells_build(_p(pos), _p(cell), _p(crystal_ptr), _p(node_crystal), N, B, float(radius),
                                                 _p(reps_arg), stride, _p(ws), st), "nlist_cells_build")
        _lib.check(lib.cartnet_nlist_cells_count(_p(pos), _p(cell), _p(crystal_ptr), _p(node_crystal), N, B, float(radius), r2,
                                                 _p(reps_arg), stride, _p(ws), _p(row_count), st), "nlist_cells_count")
    else:
        _lib.check(lib.cartnet_nlist_count(_p(pos), _p(cell), _p(crystal_ptr), _p(node_crystal), N, float(radius), r2,
                                           _p(reps_arg), stride, _p(row_count), st), "nlist_count")
    row_ptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    _lib.check(lib.cartnet_exclusive_scan_i32(_p(row_count), N, _p(row_ptr), st), "exclusive_scan")
    E = int(row_ptr[-1].item()) if N > 0 else 0          # the one host sync: output sizes
    out = {
        "edge_index": torch.empty(2, E, dtype=torch.int64, device=dev),
        "unit_cell": torch.empty(E, 3, dtype=torch.float32, device=dev),
        "dist": torch.empty(E, dtype=torch.float32, device=dev),
        "direction": torch.empty(E, 3, dtype=torch.float32, device=dev),
        "row_ptr": row_ptr, "reps": reps, "reps_max": reps_max,
    }
    if want_cart:
        out["cart_dist"] = torch.empty(E, dtype=torch.float32, device=dev)
        out["cart_dir"] = torch.empty(E, 3, dtype=torch.float32, device=dev)
    if want_i32:
        out["src32"] = torch.empty(E, dtype=torch.int32, device=dev)
        out["dst32"] = torch.empty(E, dtype=torch.int32, device=dev)
    if ws is not None:
        _lib.check(lib.cartnet_nlist_cells_fill(
            _p(pos), _p(cell), _p(crystal_ptr), _p(node_crystal), N, B, float(radius), r2, _p(reps_arg), stride, _p(ws), _p(row_ptr),
            _p(out["edge_index"]), E, _p(out["unit_cell"]), _p(out["dist"]), _p(out["direction"]),
            _p(out.get("cart_dist")), _p(out.get("cart_dir")), _p(out.get("src32")), _p(out.get("dst32")), st), "nlist_cells_fill")
    else:
        _lib.check(lib.cartnet_nlist_fill(
            _p(pos), _p(cell), _p(crystal_ptr), _p(node_crystal), N, float(radius), r2, _p(reps_arg), stride, _p(row_ptr),
            _p(out["edge_index"]), E, _p(out["unit_cell"]), _p(out["dist"]), _p(out["direction"]),
            _p(out.get("cart_dist")), _p(out.get("cart_dir")), _p(out.get("src32")), _p(out.get("dst32")), st), "nlist_fill")
    return out


def nlist_knn_mask(direction, row_ptr, num_nodes: int, threshold: int, tolerance: float = 0.01, strict: bool = False):
    """Boolean keep-mask [E] of the kNN neighbour cap (dataset/utils.py:240-360) for a dst-sorted graph."""
    lib = _lib.load()
    direction = _req(direction.contiguous(), torch.float32, "direction")
    E = int(direction.shape[0])
    keep = torch.empty(E, dtype=torch.uint8, device=direction.device)
    scratch = torch.empty(max(E, 1), dtype=torch.float32, device=direction.device)
    counts = torch.empty(num_nodes, dtype=torch.int32, device=direction.device)
    _lib.check(lib.cartnet_nlist_knn_mask(_p(direction), _p(row_ptr), num_nodes, int(threshold), float(tolerance), int(bool(strict)),
                                          _p(scratch), _p(keep), _p(counts), _stream()), "nlist_knn_mask")
    return keep.bool(), counts


@dataclass
class GraphPlan:
    """int32 indexing for the layer kernels; built once per batch and reused by all layers."""
    num_nodes: int
    num_edges: int
    src32: torch.Tensor
    dst32: torch.Tensor
    row_ptr: torch.Tensor            # dst CSR (edges are dst-sorted)
    col_ptr: torch.Tensor            # src CSR
    perm_src: torch.Tensor           # edge ids grouped by src
    perm_dst: Optional[torch.Tensor] = None   # only when the caller's edges were NOT dst-sorted
    key: tuple = ()


def graph_plan(edge_index: torch.Tensor, num_nodes: int, assume_dst_sorted: bool = False) -> GraphPlan:
    """assume_dst_sorted=True (a promise of the data pipeline: every graph produced by radius_graph_pbc is
    dst-sorted with in-range indices, SURVEY.md §4 property 3) skips the one device->host read of the check flags,
    so building the plan does not synchronise the stream."""
    lib = _lib.load()
    edge_index = _req(edge_index.contiguous(), torch.int64, "edge_index")
    dev = edge_index.device
    E = int(edge_index.shape[1])
    st = _stream()
    src32 = torch.empty(E, dtype=torch.int32, device=dev)
    dst32 = torch.empty(E, dtype=torch.int32, device=dev)
    flags = torch.empty(2, dtype=torch.int32, device=dev)
    _lib.check(lib.cartnet_graph_split(_p(edge_index), E, num_nodes, _p(src32), _p(dst32), _p(flags), st), "graph_split")
    unsorted, oob = (0, 0) if assume_dst_sorted else (int(v) for v in flags.tolist())   # host sync, once per batch
    if oob:
        raise IndexError("cartnet_b200: edge_index has entries outside [0, %d)" % num_nodes)
    cursor = torch.empty(max(num_nodes, 1), dtype=torch.int32, device=dev)

    def csr(keys):
        ptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
        perm = torch.empty(E, dtype=torch.int32, device=dev)
        _lib.check(lib.cartnet_graph_csr(_p(keys), E, num_nodes, _p(ptr), _p(perm), _p(cursor), st), "graph_csr")
        return ptr, perm

    perm_dst = None
    row_ptr, perm_d = csr(dst32)
    if unsorted:   # bring the edges into dst-sorted order; callers permute their edge tensors with perm_dst
        perm_dst = perm_d.to(torch.int64)
        src32 = src32[perm_dst].contiguous()
        dst32 = dst32[perm_dst].contiguous()
    col_ptr, perm_src = csr(src32)
    return GraphPlan(num_nodes, E, src32, dst32, row_ptr, col_ptr, perm_src, perm_dst)


# ----------------------------------------------------------------------------- featuriser
def edge_features(cart_dist, cart_dir, means, betas, cutoff_upper: float, invariant: bool, ld: int, prec: int):
    lib = _lib.load()
    cart_dist = _req(cart_dist.contiguous(), torch.float32, "cart_dist")
    if cart_dist.dim() != 1:
        raise ValueError("cart_dist must be 1-D [E] (cartnet.py:241 unsqueezes it)")
    E = int(cart_dist.shape[0])
    if not invariant:
        cart_dir = _req(cart_dir.contiguous(), torch.float32, "cart_dir")
    feat = torch.empty(E, ld, dtype=t_dtype(prec), device=cart_dist.device)
    _lib.check(lib.cartnet_edge_features(_p(cart_dist), None if invariant else _p(cart_dir), _p(means), _p(betas),
                                         int(means.numel()), float(cutoff_upper), int(bool(invariant)), E, _p(feat),
                                         ld, prec, _stream()), "edge_features")
    return feat


# ----------------------------------------------------------------------------- GEMMs
def gemm(prec: int, A, B, *, bias=None, gather0=None, gidx0=None, gather1=None, gidx1=None,
         z_out=None, act: int = ACT_NONE, z_in=None, resid=None, out_f32=None, out_t=None):
    """C[M,N] = A[M,K] @ B[N,K]^T with the fused epilogue of `cartnet_gemm_t`. Outputs are the
    caller-provided (possibly strided) views z_out / out_f32 / out_t."""
    lib = _lib.load()
    T = t_dtype(prec)
    _req(A, T, "A"); _req(B, T, "B")
    M, K = int(A.shape[0]), int(A.shape[1])
    N = int(B.shape[0])
    if int(B.shape[1]) != K:
        raise ValueError("gemm: K mismatch %d vs %d" % (K, int(B.shape[1])))
    d = _lib.GemmDesc()
    d.prec, d.M, d.N, d.K = prec, M, N, K
    d.A, d.lda, d.B, d.ldb = _p(A), _ld2(A), _p(B), _ld2(B)
    d.bias = _p(_req(bias, torch.float32, "bias")) if bias is not None else None
    if gather0 is not None:
        _req(gather0, T, "gather0"); _req(gidx0, torch.int32, "gidx0")
        d.gather0, d.gidx0, d.ldg = _p(gather0), _p(gidx0), _ld2(gather0)
    if gather1 is not None:
        _req(gather1, T, "gather1"); _req(gidx1, torch.int32, "gidx1")
        if gather0 is not None and _ld2(gather1) != d.ldg:
            raise ValueError("gemm: gather0/gather1 must share a leading dimension")
        d.gather1, d.gidx1, d.ldg = _p(gather1), _p(gidx1), _ld2(gather1)
    if z_out is not None:
        _req(z_out, z_dtype(prec), "z_out"); d.z_out, d.ldz = _p(z_out), _ld2(z_out)
    d.act = act
    if z_in is not None:
        _req(z_in, z_dtype(prec), "z_in"); d.z_in, d.ldzin = _p(z_in), _ld2(z_in)
    if resid is not None:
        _req(resid, torch.float32, "resid"); d.resid, d.ldr = _p(resid), _ld2(resid)
    if out_f32 is not None:
        _req(out_f32, torch.float32, "out_f32"); d.out_f32, d.ldo = _p(out_f32), _ld2(out_f32)
    if out_t is not None:
        _req(out_t, T, "out_t"); d.out_t, d.ldt = _p(out_t), _ld2(out_t)
    for nm, t in (("z_out", z_out), ("z_in", z_in), ("resid", resid), ("out_f32", out_f32), ("out_t", out_t)):
        if t is not None and (int(t.shape[0]) != M or int(t.shape[1]) != N):
            raise ValueError("gemm: %s has shape %s, expected [%d,%d]" % (nm, tuple(t.shape), M, N))
    _lib.check(lib.cartnet_gemm(C.byref(d), _stream()), "gemm")


def gemm_colstats(prec: int, A, B, bias, out_t, running_mean=None, running_var=None, momentum: float = 0.1, shift=None):
    """out_t = (T)(A @ B^T + bias) together with the column mean / biased variance of that output over the M rows
    (BatchNorm batch statistics; running buffers updated in place, `shift` as in colstats). Tensor-core modes
    accumulate the sums in the GEMM epilogue, fp32 mode runs the statistics pass after the GEMM."""
    lib = _lib.load()
    T = t_dtype(prec)
    _req(A, T, "A"); _req(B, T, "B"); _req(out_t, T, "out_t"); _req(bias, torch.float32, "bias")
    M, K, N = int(A.shape[0]), int(A.shape[1]), int(B.shape[0])
    if int(B.shape[1]) != K or tuple(out_t.shape) != (M, N):
        raise ValueError("gemm_colstats: shape mismatch")
    d = _lib.GemmDesc()
    d.prec, d.M, d.N, d.K = prec, M, N, K
    d.A, d.lda, d.B, d.ldb = _p(A), _ld2(A), _p(B), _ld2(B)
    d.bias, d.act = _p(bias), ACT_NONE
    d.out_t, d.ldt = _p(out_t), _ld2(out_t)
    mean = torch.empty(N, dtype=torch.float32, device=A.device)
    var = torch.empty(N, dtype=torch.float32, device=A.device)
    part = _partial(A.device, int(lib.cartnet_colstats_workspace(N)))
    _lib.check(lib.cartnet_gemm_colstats(C.byref(d), _p(shift), _p(mean), _p(var), _p(running_mean), _p(running_var),
                                         float(momentum), _p(part), _stream()), "gemm_colstats")
    return mean, var


def gemm_tn(prec: int, A, B, out_blocks=None) -> torch.Tensor:
    """C[M,N] fp32 = A[K,M]^T @ B[K,N] (deterministic split-K). With out_blocks = [v_0 .. v_{b-1}] (b <= 4 fp32 views
    of shape [M/b, N] sharing one row pitch) row block i of C is written into v_i instead of a fresh tensor."""
    lib = _lib.load()
    T = t_dtype(prec)
    _req(A, T, "A"); _req(B, T, "B")
    K, M, N = int(A.shape[0]), int(A.shape[1]), int(B.shape[1])
    if int(B.shape[0]) != K:
        raise ValueError("gemm_tn: K mismatch")
    nbytes = int(lib.cartnet_gemm_tn_workspace(prec, M, N, K))
    ws = _workspace(A.device, nbytes)
    if out_blocks is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
        _lib.check(lib.cartnet_gemm_tn(prec, M, N, K, _p(A), _ld2(A), _p(B), _ld2(B), _p(out), N, _p(ws),
                                       ws.numel() * 4, _stream()), "gemm_tn")
        return out
    nb = len(out_blocks)
    for v in out_blocks:
        _req(v, torch.float32, "out_blocks[i]")
        if tuple(v.shape) != (M // nb, N) or _ld2(v) != _ld2(out_blocks[0]):
            raise ValueError("gemm_tn: out_blocks must be %d views of shape [%d,%d] with one row pitch" % (nb, M // nb, N))
    ptrs = (C.c_void_p * nb)(*[v.data_ptr() for v in out_blocks])
    _lib.check(lib.cartnet_gemm_tn_blocks(prec, M, N, K, _p(A), _ld2(A), _p(B), _ld2(B), ptrs, nb, _ld2(out_blocks[0]),
                                          _p(ws), ws.numel() * 4, _stream()), "gemm_tn_blocks")
    return out_blocks


# ----------------------------------------------------------------------------- reductions
def colstats(x, running_mean=None, running_var=None, momentum: float = 0.1, shift=None, prec: int = PREC_FP32):
    """Per-column mean / biased variance over rows of x (fp32, or T of `prec`); updates running stats in place
    (train-mode BN). `shift` [C]: x was stored centred (x = true - shift); only the running-mean update sees it."""
    lib = _lib.load()
    is_t = 0 if x.dtype == torch.float32 and prec not in (PREC_TF32, PREC_BF16X3) else 1
    _req(x, t_dtype(prec) if is_t else torch.float32, "x")
    rows, Cc = int(x.shape[0]), int(x.shape[1])
    mean = torch.empty(Cc, dtype=torch.float32, device=x.device)
    var = torch.empty(Cc, dtype=torch.float32, device=x.device)
    part = _partial(x.device, int(lib.cartnet_colstats_workspace(Cc)))
    _lib.check(lib.cartnet_colstats(_p(x), is_t, prec, rows, Cc, _ld2(x), _p(shift), _p(mean), _p(var), _p(running_mean),
                                    _p(running_var), float(momentum), _p(part), _stream()), "colstats")
    return mean, var


def gate_center(H_g, G2, bg2, running_mean, training: bool, prec: int):
    """(bias_c, center) for the centred gate GEMM g - center = H_g G2^T + bias_c; see cartnet_gate_center."""
    lib = _lib.load()
    _req(H_g, t_dtype(prec), "H_g"); _req(G2, torch.float32, "G2")
    E, D = int(H_g.shape[0]), int(H_g.shape[1])
    dev = H_g.device
    out = torch.empty(3, D, dtype=torch.float32, device=dev)
    part = _partial(dev, int(lib.cartnet_colstats_workspace(D)))
    _lib.check(lib.cartnet_gate_center(_p(H_g), _ld2(H_g), E, D, _p(G2.contiguous()), _p(bg2), _p(running_mean),
                                       int(bool(training)), prec, _p(out[0]), _p(out[1]), _p(out[2]), _p(part), _stream()),
               "gate_center")
    return out[0], out[1]


def colsum(x, prec: int) -> torch.Tensor:
    """column sums of a T-typed tensor of mode `prec` (fp32 words in the fp32 / tf32 modes)"""
    lib = _lib.load()
    is_t = 1 if (x.dtype != torch.float32 or prec == PREC_BF16X3) else 0
    _req(x, torch.float32 if not is_t else t_dtype(prec), "x")
    rows, Cc = int(x.shape[0]), int(x.shape[1])
    out = torch.empty(Cc, dtype=torch.float32, device=x.device)
    part = _partial(x.device, int(lib.cartnet_colstats_workspace(Cc)))
    _lib.check(lib.cartnet_colsum(_p(x), is_t, prec, rows, Cc, _ld2(x), _p(out), _p(part), _stream()), "colsum")
    return out


# ----------------------------------------------------------------------------- layer passes
def edge_gate_aggregate(g, s, e, dist, row_ptr, num_nodes: int, bn_mean, bn_var, bn_w, bn_b, radius: float,
                        use_envelope: bool, prec: int, want_shadow: bool, want_gn: bool = True):
    """g, s: T (g possibly stored centred, bn_mean = mean of the stored values, None = 0). Returns e_out (fp32), its T
    shadow, m, and gn_t = (g-mean)*rstd in T (what the backward pass reads instead of g)."""
    lib = _lib.load()
    T = t_dtype(prec)
    for nm, t, dt in (("g", g, T), ("s", s, T), ("e", e, torch.float32)):
        _req(t, dt, nm)
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % nm)
    E, D = int(e.shape[0]), int(e.shape[1])
    e_out = torch.empty_like(e)
    e_out_t = torch.empty(E, D, dtype=T, device=e.device) if (want_shadow and needs_shadow(prec)) else None
    gn_t = torch.empty(E, D, dtype=T, device=e.device) if want_gn else None
    m = torch.empty(num_nodes, D, dtype=torch.float32, device=e.device)
    _lib.check(lib.cartnet_edge_gate_aggregate(
        _p(g), _p(s), _p(e), _p(dist), _p(row_ptr), num_nodes, E, D, _p(bn_mean), _p(bn_var), _p(bn_w), _p(bn_b),
        EPS_BN, float(radius), int(bool(use_envelope)), _p(e_out), _p(e_out_t), _p(gn_t), prec, _p(m), _stream()),
        "edge_gate_aggregate")
    return e_out, (e_out_t if e_out_t is not None else (e_out if not needs_shadow(prec) else None)), m, gn_t


def node_update(m, x, bn_mean, bn_var, bn_w, bn_b, prec: int, want_shadow: bool):
    lib = _lib.load()
    _req(m, torch.float32, "m"); _req(x, torch.float32, "x")
    x = x.contiguous()
    N, D = int(x.shape[0]), int(x.shape[1])
    x_out = torch.empty_like(x)
    x_out_t = torch.empty(N, D, dtype=t_dtype(prec), device=x.device) if (want_shadow and needs_shadow(prec)) else None
    _lib.check(lib.cartnet_node_update(_p(m), _p(x), N, D, _p(bn_mean), _p(bn_var), _p(bn_w), _p(bn_b), EPS_BN,
                                       _p(x_out), _p(x_out_t), prec, _stream()), "node_update")
    return x_out, (x_out_t if x_out_t is not None else (x_out if not needs_shadow(prec) else None))


def node_update_bwd(dx_out, m, bn_mean, bn_var, bn_w, bn_b, training: bool):
    """Returns dm [N,D] and sums [2D] (sum dy | sum dy*yhat) = (d bias | d weight) of norm2."""
    lib = _lib.load()
    dx_out = _req(dx_out.contiguous(), torch.float32, "dx_out")
    N, D = int(m.shape[0]), int(m.shape[1])
    sums = torch.empty(2 * D, dtype=torch.float32, device=m.device)
    part = _partial(m.device, int(lib.cartnet_colstats_workspace(D)))
    st = _stream()
    _lib.check(lib.cartnet_node_update_bwd_reduce(_p(dx_out), _p(m), N, D, _p(bn_mean), _p(bn_var), _p(bn_w), _p(bn_b),
                                                  EPS_BN, _p(sums), _p(part), st), "node_update_bwd_reduce")
    dm = torch.empty_like(m)
    _lib.check(lib.cartnet_node_update_bwd_apply(_p(dx_out), _p(m), N, D, _p(bn_mean), _p(bn_var), _p(bn_w), _p(bn_b),
                                                 EPS_BN, _p(sums), int(bool(training)), _p(dm), st), "node_update_bwd_apply")
    return dm, sums


def edge_gate_bwd(gn_t, s_t, dist, dst32, de_out, dm, bn_var, bn_w, bn_b, radius: float, use_envelope: bool,
                  training: bool, prec: int, g_mean=None, input_is_g: bool = False):
    """gn_t, s_t: the T tensors saved by edge_gate_aggregate / the MLP_aggr GEMM. de_out may be None (= zero).
    input_is_g: `gn_t` is the stored, centred pre-activation g itself (no normalised copy was written forward);
    gn = (g - g_mean) * rsqrt(bn_var + eps) is formed inside the kernels (g_mean None = 0).
    Returns ds_t, dg_t (T, [E,D]) and sums [3D]: sum dghat (= d bias of the edge BatchNorm) | sum dghat*gn
    (= d weight) | sum ds (= d bias of MLP_aggr[2])."""
    lib = _lib.load()
    T = t_dtype(prec)
    _req(gn_t, T, "gn_t"); _req(s_t, T, "s_t")
    if de_out is not None:
        de_out = _req(de_out.contiguous(), torch.float32, "de_out")
    E, D = int(gn_t.shape[0]), int(gn_t.shape[1])
    dev = gn_t.device
    ds_t = torch.empty(E, D, dtype=T, device=dev)
    dg_t = torch.empty(E, D, dtype=T, device=dev)
    dghat_t = torch.empty(E, D, dtype=T, device=dev)
    sums = torch.empty(3 * D, dtype=torch.float32, device=dev)
    part = _partial(dev, int(lib.cartnet_colstats_workspace(D)))
    st = _stream()
    _lib.check(lib.cartnet_edge_gate_bwd_reduce(
        _p(gn_t), _p(s_t), _p(dist), _p(dst32), _p(de_out), _p(dm), E, D, _p(bn_w), _p(bn_b),
        float(radius), int(bool(use_envelope)), _p(ds_t), _p(dghat_t), prec, _p(sums), _p(part),
        _p(g_mean) if input_is_g else None, _p(bn_var) if input_is_g else None, EPS_BN, st), "edge_gate_bwd_reduce")
    _lib.check(lib.cartnet_edge_gate_bwd_apply(_p(gn_t), _p(dghat_t), E, D, _p(bn_var), _p(bn_w), EPS_BN,
                                               _p(sums), int(bool(training)), _p(dg_t), prec,
                                               _p(g_mean) if input_is_g else None, int(bool(input_is_g)), st), "edge_gate_bwd_apply")
    return ds_t, dg_t, sums


def segment_sum(x, ptr, perm, num_nodes: int, out, prec: int):
    """out[n,:] = sum over CSR row n of x[perm[k],:]; `out` is a caller-provided [N,C] view (T or fp32)."""
    lib = _lib.load()
    _req(x, t_dtype(prec), "x")
    Cc = int(x.shape[1])
    out_is_t = 0 if (out.dtype == torch.float32 and not f32_storage(prec)) else 1
    _lib.check(lib.cartnet_segment_sum(_p(x), _ld2(x), _p(ptr), _p(perm), num_nodes, Cc, _p(out), _ld2(out),
                                       out_is_t, prec, _stream()), "segment_sum")
    return out


def segment_sum_pair(x, row_ptr, col_ptr, perm_src, num_nodes: int, out, prec: int):
    """out[:, :C] = dst-CSR segment sum of x, out[:, C:2C] = src-CSR segment sum (through perm_src), one launch;
    `out` is a caller-provided [N, 2C] tensor (T or fp32)."""
    lib = _lib.load()
    _req(x, t_dtype(prec), "x")
    Cc = int(x.shape[1])
    if tuple(out.shape) != (num_nodes, 2 * Cc):
        raise ValueError("segment_sum_pair: out must be [N, 2C]")
    out_is_t = 0 if (out.dtype == torch.float32 and not f32_storage(prec)) else 1
    _lib.check(lib.cartnet_segment_sum_pair(_p(x), _ld2(x), _p(row_ptr), _p(col_ptr), _p(perm_src), num_nodes, Cc, _p(out),
                                            _ld2(out), out_is_t, prec, _stream()), "segment_sum_pair")
    return out


def dsilu_mul(dy, z, prec: int, want_colsum: bool = False):
    """y = (T)(dy * silu'(z)); with want_colsum also the column sums of y (fp32 [C]) from the same pass."""
    lib = _lib.load()
    _req(dy, torch.float32, "dy"); _req(z, z_dtype(prec), "z")
    rows, Cc = int(z.shape[0]), int(z.shape[1])
    y = torch.empty(rows, Cc, dtype=t_dtype(prec), device=z.device)
    tpr = Cc // 4
    if want_colsum and Cc % 4 == 0 and 1 <= tpr <= 256 and tpr & (tpr - 1) == 0:
        cs = torch.empty(Cc, dtype=torch.float32, device=z.device)
        part = _partial(z.device, int(lib.cartnet_colstats_workspace(Cc)))
        _lib.check(lib.cartnet_dsilu_mul(_p(dy), _ld2(dy), _p(z), _ld2(z), _p(y), Cc, rows, Cc, prec, _p(cs), _p(part), _stream()), "dsilu_mul")
        return y, cs
    _lib.check(lib.cartnet_dsilu_mul(_p(dy), _ld2(dy), _p(z), _ld2(z), _p(y), Cc, rows, Cc, prec, None, None, _stream()), "dsilu_mul")
    return (y, colsum(y, prec)) if want_colsum else y


def cast(x, prec: int) -> torch.Tensor:
    """fp32 -> T operand copy: bf16, or fp32 rounded to tf32; the identity object in fp32 mode."""
    if not needs_shadow(prec):
        return x
    lib = _lib.load()
    _req(x, torch.float32, "x")
    rows, Cc = int(x.shape[0]), int(x.shape[1])
    y = torch.empty(rows, Cc, dtype=t_dtype(prec), device=x.device)
    _lib.check(lib.cartnet_cast_rows(_p(x), _ld2(x), _p(y), Cc, rows, Cc, prec, _stream()), "cast_rows")
    return y


def uncast(x_t, prec: int) -> torch.Tensor:
    """T -> fp32 values (exact): the only way to read a bf16x3 buffer outside the library's kernels."""
    lib = _lib.load()
    _req(x_t, t_dtype(prec), "x_t")
    rows, Cc = int(x_t.shape[0]), int(x_t.shape[1])
    y = torch.empty(rows, Cc, dtype=torch.float32, device=x_t.device)
    _lib.check(lib.cartnet_uncast_rows(_p(x_t), _ld2(x_t), _p(y), Cc, rows, Cc, prec, _stream()), "uncast_rows")
    return y


# ----------------------------------------------------------------------------- Cholesky head tail
def cholesky_head_fwd(h, W1, b1):
    """h [n, Dh] fp32 -> (U [n,3,3] = L^T L, p6 [n,6] = h W1^T + b1) in one launch (cartnet.py:293-303)."""
    lib = _lib.load()
    _req(h, torch.float32, "h"); _req(W1, torch.float32, "W1"); _req(b1, torch.float32, "b1", rowmajor=False)
    n, Dh = int(h.shape[0]), int(h.shape[1])
    p6 = torch.empty(n, 6, dtype=torch.float32, device=h.device)
    U = torch.empty(n, 3, 3, dtype=torch.float32, device=h.device)
    _lib.check(lib.cartnet_cholesky_head_fwd(_p(h), _ld2(h), _p(W1), _p(b1), n, Dh, _p(p6), _p(U), _stream()), "cholesky_head_fwd")
    return U, p6


def cholesky_head_bwd(dU, h, p6, W1):
    """-> (dh [n, Dh], dW1 [6, Dh], db1 [6]); deterministic reduction over atoms."""
    lib = _lib.load()
    _req(dU, torch.float32, "dU", rowmajor=False); _req(h, torch.float32, "h"); _req(p6, torch.float32, "p6")
    n, Dh = int(h.shape[0]), int(h.shape[1])
    dh = torch.empty(n, Dh, dtype=torch.float32, device=h.device)
    dW1 = torch.empty(6, Dh, dtype=torch.float32, device=h.device)
    db1 = torch.empty(6, dtype=torch.float32, device=h.device)
    part = _workspace(h.device, int(lib.cartnet_cholesky_head_workspace(n, Dh)))
    _lib.check(lib.cartnet_cholesky_head_bwd(_p(dU), _p(h), _ld2(h), _p(p6), _p(W1), n, Dh, _p(dh), Dh, _p(dW1), _p(db1),
                                             _p(part), _stream()), "cholesky_head_bwd")
    return dh, dW1, db1


# ----------------------------------------------------------------------------- loss pair
def loss_l1_mse(pred, true):
    """(2,) fp32 = [mean |pred - true|, mean (pred - true)^2] in one launch (train/metrics.py:15-28)."""
    lib = _lib.load()
    _req(pred, torch.float32, "pred"); _req(true, torch.float32, "true")
    out = torch.empty(2, dtype=torch.float32, device=pred.device)
    _lib.check(lib.cartnet_loss_l1_mse(_p(pred), _p(true), int(pred.numel()), _p(out), _stream()), "loss_l1_mse")
    return out


def loss_l1_mse_bwd(pred, true, dmae, dmse):
    """dpred for upstream gradients dmae / dmse (0-dim or 1-element DEVICE tensors, or None)."""
    lib = _lib.load()
    dpred = torch.empty_like(pred)
    _lib.check(lib.cartnet_loss_l1_mse_bwd(_p(pred), _p(true), int(pred.numel()), _p(dmae), _p(dmse), _p(dpred), _stream()), "loss_l1_mse_bwd")
    return dpred
