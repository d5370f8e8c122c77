"""Tensor-level wrappers over the C ABI (``include/laff_b200.h``).

PyTorch is used for device memory and streams only; every computation below is a hand-written sm_100a kernel reached
through ctypes.  All functions require CUDA tensors and raise :class:`LaffError` on failure — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import os

import numpy as np
import torch

from . import _capi
from ._capi import (BF16, F16, F32, FUSE_MAX_FC, FUSE_MAX_TILED, MAX_TOPK, MAX_TOPK_DENSE, FuseDesc, LaffError,
                    PoolDesc)

_DT = {torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32}
_TORCH_DT = {F16: torch.float16, BF16: torch.bfloat16, F32: torch.float32, "fp16": torch.float16,
             "bf16": torch.bfloat16, "fp32": torch.float32}


def torch_dtype(d) -> torch.dtype:
    return d if isinstance(d, torch.dtype) else _TORCH_DT[d]


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise LaffError("laff_b200 ops need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D view whose last dim is contiguous (row pitch = stride(0))."""
    if t.dim() != 2:
        t = t.reshape(t.shape[0], -1)
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def set_tuning(cta_group: int = 0, chunk_tiles: int = 0, m_group: int = 0) -> None:
    _capi.call("laff_set_tuning", cta_group, chunk_tiles, m_group)


def get_tuning() -> tuple[int, int, int]:
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _capi.call("laff_get_tuning", C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def set_fuse_variant(cta_group: int = 0) -> None:
    """0 = per-call choice, 1 / 2 = force the cta_group::1 / ::2 single-kernel fusion (include/laff_b200.h)."""
    _capi.call("laff_set_fuse_variant", cta_group)


def get_fuse_variant() -> int:
    return int(_capi.lib().laff_get_fuse_variant())


# ----------------------------------------------------------------------------------------------------------------
# operand preparation
# ----------------------------------------------------------------------------------------------------------------
def l2norm_quantize(x: torch.Tensor, heads: int, out_dtype=torch.bfloat16, eps: float = 1e-13 + 1e-14,
                    normalise: bool = True) -> torch.Tensor:
    """loss.l2norm per head (loss.py:8-13) then rounding to ``out_dtype``. x [rows, heads*dh] (or [rows, heads, dh])."""
    _need_cuda(x)
    shape = x.shape
    x2 = _rowmajor(x.float() if x.dtype != torch.float32 else x)
    rows, D = x2.shape
    if D % heads:
        raise LaffError("feature dim %d not divisible by heads %d" % (D, heads))
    out_dtype = torch_dtype(out_dtype)
    out = torch.empty((rows, D), dtype=out_dtype, device=x.device)
    if rows:
        _capi.call("laff_l2norm_quantize", _ptr(x2), rows, heads, D // heads, x2.stride(0),
                   float(eps) if normalise else -1.0, _DT[out_dtype], _ptr(out), out.stride(0), _stream(x))
    return out.reshape(shape)


def cast_pad_16(x: torch.Tensor, dtype=torch.bfloat16, multiple: int = 8, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> 16-bit [rows, cols_pad] with zero padding so the row pitch is TMA-legal.  `out`: optional
    16-bit scratch of at least [rows, cols_pad] to write into (its first `rows` rows are returned)."""
    _need_cuda(x)
    x2 = _rowmajor(x.float() if x.dtype != torch.float32 else x)
    rows, cols = x2.shape
    cols_pad = (cols + multiple - 1) // multiple * multiple
    dtype = torch_dtype(dtype)
    if out is None:
        out = torch.empty((rows, cols_pad), dtype=dtype, device=x.device)
    else:
        if out.dtype != dtype or out.dim() != 2 or out.shape[0] < rows or out.shape[1] != cols_pad or out.stride(1) != 1:
            raise LaffError("cast_pad_16: out must be a %s [>=%d, %d] row-major tensor" % (dtype, rows, cols_pad))
        out = out[:rows]
    if rows:
        _capi.call("laff_cast_pad_16", _ptr(x2), rows, cols, x2.stride(0), _DT[dtype], _ptr(out), cols_pad,
                   out.stride(0), _stream(x))
    return out


def split3_16(x: torch.Tensor, side: int, dtype=torch.bfloat16, multiple: int = 8) -> torch.Tensor:
    """3-term split operands ([hi|lo|hi] for side 0, [hi|hi|lo] for side 1) for near-fp32 tensor-core products."""
    _need_cuda(x)
    x2 = _rowmajor(x.float() if x.dtype != torch.float32 else x)
    rows, cols = x2.shape
    cols_pad = (cols + multiple - 1) // multiple * multiple
    dtype = torch_dtype(dtype)
    out = torch.empty((rows, 3 * cols_pad), dtype=dtype, device=x.device)
    if rows:
        _capi.call("laff_split3_16", _ptr(x2), rows, cols, x2.stride(0), side, _DT[dtype], _ptr(out), cols_pad,
                   out.stride(0), _stream(x))
    return out


# ----------------------------------------------------------------------------------------------------------------
# similarity / ranking
# ----------------------------------------------------------------------------------------------------------------
def _check_operands(q: torch.Tensor, g: torch.Tensor):
    _need_cuda(q, g)
    if q.dtype != g.dtype or q.dtype not in (torch.float16, torch.bfloat16):
        raise LaffError("similarity operands must both be fp16 or bf16 (got %s, %s)" % (q.dtype, g.dtype))
    q, g = _rowmajor(q), _rowmajor(g)
    if q.shape[1] != g.shape[1]:
        raise LaffError("operand K mismatch: %d vs %d" % (q.shape[1], g.shape[1]))
    return q, g


def sim_dense(q: torch.Tensor, g: torch.Tensor, scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i, j] = scale * <q_i, g_j>, fp32. q [Q, D], g [V, D] 16-bit."""
    q, g = _check_operands(q, g)
    Q, D = q.shape
    V = g.shape[0]
    if out is None:
        out = torch.empty((Q, V), dtype=torch.float32, device=q.device)
    if Q and V:
        _capi.call("laff_sim_dense", _ptr(q), _ptr(g), Q, V, D, q.stride(0), g.stride(0), _DT[q.dtype], float(scale),
                   _ptr(out), out.stride(0), _stream(q))
    return out


def sim_collect(q: torch.Tensor, g: torch.Tensor, thr: torch.Tensor, cap: int, scale: float = 1.0, col_offset: int = 0,
                sgt_raw: Optional[torch.Tensor] = None, gt_global: Optional[torch.Tensor] = None):
    """Candidates of long ranked lists without the dense matrix (laff_sim_collect): every score scale * <q_i, g_j> >=
    thr[i], unordered.  -> (count int32 [Q], cand_val fp32 [Q, cap], cand_idx int32 [Q, cap] global indices, -inf / -1
    in unused slots).  count[i] > cap means query i's list was truncated.  With sgt_raw / gt_global (as for
    sim_rank_topk) the same sweep also counts the ground truth's local rank: a fourth result, int32 [Q]."""
    q, g = _check_operands(q, g)
    Q, D = q.shape
    _need_cuda(thr)
    thr = thr.to(torch.float32).contiguous()
    if thr.numel() != Q:
        raise LaffError("sim_collect: thr must have one entry per query")
    count = torch.empty(Q, dtype=torch.int32, device=q.device)
    cv = torch.empty((Q, cap), dtype=torch.float32, device=q.device)
    ci = torch.empty((Q, cap), dtype=torch.int32, device=q.device)
    if sgt_raw is not None:
        _need_cuda(sgt_raw, gt_global)
        sgt_raw, gt_global = sgt_raw.float().contiguous(), gt_global.to(torch.int32).contiguous()
        rank = torch.empty(Q, dtype=torch.int32, device=q.device)
        _capi.call("laff_sim_collect_rank", _ptr(q), _ptr(g), Q, g.shape[0], D, q.stride(0), g.stride(0), _DT[q.dtype], float(scale),
                   _ptr(thr), int(col_offset), int(cap), _ptr(count), _ptr(cv), _ptr(ci), _ptr(sgt_raw), _ptr(gt_global), _ptr(rank),
                   _stream(q))
        return count, cv, ci, rank
    _capi.call("laff_sim_collect", _ptr(q), _ptr(g), Q, g.shape[0], D, q.stride(0), g.stride(0), _DT[q.dtype], float(scale),
               _ptr(thr), int(col_offset), int(cap), _ptr(count), _ptr(cv), _ptr(ci), _stream(q))
    return count, cv, ci


def sim_gt_scores(q: torch.Tensor, g: torch.Tensor, gt_local: torch.Tensor) -> torch.Tensor:
    """Raw (unscaled) accumulator of <q_i, g_{gt_local[i]}>; 0 where gt_local[i] < 0."""
    q, g = _check_operands(q, g)
    _need_cuda(gt_local)
    Q, D = q.shape
    gt_local = gt_local.to(torch.int32).contiguous()
    sgt = torch.zeros(Q, dtype=torch.float32, device=q.device)
    nbytes = _capi.lib().laff_sim_gt_workspace_bytes(Q, D)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    _capi.call("laff_sim_gt_scores", _ptr(q), _ptr(g), Q, g.shape[0], D, q.stride(0), g.stride(0), _DT[q.dtype],
               _ptr(gt_local), _ptr(sgt), _ptr(ws), nbytes, _stream(q))
    return sgt


def sim_rank_topk(q: torch.Tensor, g: torch.Tensor, sgt_raw: torch.Tensor, gt_global: torch.Tensor, k: int,
                  scale: float = 1.0, col_offset: int = 0, workspace: Optional[torch.Tensor] = None):
    """One fused sweep: (count int32 [Q], topk_val fp32 [Q,k], topk_idx int32 [Q,k]) over the local gallery shard."""
    q, g = _check_operands(q, g)
    _need_cuda(sgt_raw, gt_global)
    Q, D = q.shape
    V = g.shape[0]
    gt_global = gt_global.to(torch.int32).contiguous()
    sgt_raw = sgt_raw.float().contiguous()
    count = torch.empty(Q, dtype=torch.int32, device=q.device)
    tv = torch.empty((Q, k), dtype=torch.float32, device=q.device)
    ti = torch.empty((Q, k), dtype=torch.int32, device=q.device)
    nbytes = _capi.lib().laff_sim_rank_workspace_bytes(Q, V, D)
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    _capi.call("laff_sim_rank_topk", _ptr(q), _ptr(g), Q, V, D, q.stride(0), g.stride(0), _DT[q.dtype], float(scale),
               _ptr(sgt_raw), _ptr(gt_global), int(col_offset), int(k), _ptr(count), _ptr(tv), _ptr(ti),
               _ptr(workspace), workspace.numel(), _stream(q))
    return count, tv, ti


def topk_merge(vals: torch.Tensor, idx: torch.Tensor, k_out: int, in_scale: float = 1.0):
    """vals/idx [n_lists, Q, k_in] ordered lists -> merged ([Q, k_out], [Q, k_out])."""
    _need_cuda(vals, idx)
    vals = vals.float().contiguous()
    idx = idx.to(torch.int32).contiguous()
    n_lists, Q, k_in = vals.shape
    ov = torch.empty((Q, k_out), dtype=torch.float32, device=vals.device)
    oi = torch.empty((Q, k_out), dtype=torch.int32, device=vals.device)
    _capi.call("laff_topk_merge", _ptr(vals), _ptr(idx), n_lists, Q, k_in, Q * k_in, k_out, float(in_scale), _ptr(ov),
               _ptr(oi), _stream(vals))
    return ov, oi


def rank_from_scores(scores: torch.Tensor, gt: Optional[torch.Tensor], k: int = 0):
    """Tie-rule rank / top-k of a materialised fp32 score matrix."""
    _need_cuda(scores, gt)
    scores = _rowmajor(scores.float() if scores.dtype != torch.float32 else scores)
    Q, V = scores.shape
    rank0 = None
    if gt is not None:
        gt = gt.to(torch.int32).contiguous()
        rank0 = torch.empty(Q, dtype=torch.int32, device=scores.device)
    tv = ti = None
    if k > 0:
        tv = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
        ti = torch.empty((Q, k), dtype=torch.int32, device=scores.device)
    _capi.call("laff_rank_from_scores", _ptr(scores), Q, V, scores.stride(0), _ptr(gt), int(k), _ptr(rank0), _ptr(tv),
               _ptr(ti), _stream(scores))
    return rank0, tv, ti


def topk_dense(scores: torch.Tensor, k: int, idx_in: Optional[torch.Tensor] = None, scale: float = 1.0):
    """Ranked list of every row of a dense fp32 matrix: (values [R, k], indices int32 [R, k]) ordered by the tie rule
    (score desc, index desc), 1 <= k <= 2048 (laff_topk_dense).  idx_in: int32 [R, C] indices carried by the candidates
    (-1 = empty), C <= 16384 -- merges per-shard lists.  Slots past the number of candidates hold -inf / -1."""
    _need_cuda(scores, idx_in)
    if not 1 <= k <= MAX_TOPK_DENSE:
        raise LaffError("topk_dense: k must be in [1, %d] (got %d)" % (MAX_TOPK_DENSE, k))
    scores = _rowmajor(scores.float() if scores.dtype != torch.float32 else scores)
    R, Cn = scores.shape
    ld_idx = 0
    if idx_in is not None:
        idx_in = _rowmajor(idx_in.to(torch.int32))
        if tuple(idx_in.shape) != (R, Cn):
            raise LaffError("topk_dense: idx_in must have the shape of scores")
        ld_idx = idx_in.stride(0)
    if R == 0 or Cn == 0:  # nothing to rank: empty lists
        return (torch.full((R, k), float("-inf"), dtype=torch.float32, device=scores.device),
                torch.full((R, k), -1, dtype=torch.int32, device=scores.device))
    tv = torch.empty((R, k), dtype=torch.float32, device=scores.device)
    ti = torch.empty((R, k), dtype=torch.int32, device=scores.device)
    _capi.call("laff_topk_dense", _ptr(scores), scores.stride(0), _ptr(idx_in), ld_idx, R, Cn, int(k), float(scale), _ptr(tv),
               _ptr(ti), _stream(scores))
    return tv, ti


def rank_multi_gt(scores: torch.Tensor, gt_offsets: torch.Tensor, gt_cols: torch.Tensor) -> torch.Tensor:
    """0-based tie-rule rank of every ground-truth column of every row; CSR lists gt_cols[gt_offsets[i]:gt_offsets[i+1]]
    (laff_rank_multi_gt, the video -> text direction of predictor.py:262-270)."""
    _need_cuda(scores, gt_offsets, gt_cols)
    scores = _rowmajor(scores.float() if scores.dtype != torch.float32 else scores)
    R, Cn = scores.shape
    gt_offsets = gt_offsets.to(torch.int64).contiguous()
    gt_cols = gt_cols.to(torch.int32).contiguous()
    if gt_offsets.numel() != R + 1:
        raise LaffError("rank_multi_gt: gt_offsets must have rows + 1 entries")
    rank0 = torch.empty(gt_cols.numel(), dtype=torch.int32, device=scores.device)
    _capi.call("laff_rank_multi_gt", _ptr(scores), scores.stride(0), R, Cn, _ptr(gt_offsets), _ptr(gt_cols), _ptr(rank0),
               _stream(scores))
    return rank0


def multi_gt_metrics(rank0: torch.Tensor, gt_offsets: torch.Tensor):
    """evaluation.eval from multi-ground-truth ranks: (metrics double[8], first int32 [R], ap double [R])."""
    _need_cuda(rank0, gt_offsets)
    rank0 = rank0.to(torch.int32).contiguous()
    gt_offsets = gt_offsets.to(torch.int64).contiguous()
    R = gt_offsets.numel() - 1
    first = torch.empty(R, dtype=torch.int32, device=rank0.device)
    ap = torch.empty(R, dtype=torch.float64, device=rank0.device)
    out = torch.empty(8, dtype=torch.float64, device=rank0.device)
    _capi.call("laff_multi_gt_metrics", _ptr(rank0), _ptr(gt_offsets), R, _ptr(first), _ptr(ap), _ptr(out), _stream(rank0))
    return out, first, ap


def rank_metrics(rank0: torch.Tensor) -> torch.Tensor:
    """Device tensor of 8 doubles: R@1, R@5, R@10, MedR, MeanR, MIR, mAP, Q (evaluation.py:81-89)."""
    _need_cuda(rank0)
    rank0 = rank0.to(torch.int32).contiguous()
    out = torch.empty(8, dtype=torch.float64, device=rank0.device)
    _capi.call("laff_rank_metrics", _ptr(rank0), rank0.numel(), _ptr(out), _stream(rank0))
    return out


# ----------------------------------------------------------------------------------------------------------------
# fusion
# ----------------------------------------------------------------------------------------------------------------
def bn_fold(weight, bias, running_mean, running_var, eps: float = 1e-5):
    _need_cuda(running_mean, running_var)
    D = running_mean.numel()
    scale = torch.empty(D, dtype=torch.float32, device=running_mean.device)
    shift = torch.empty_like(scale)
    f = lambda t: None if t is None else t.detach().float().contiguous()
    weight, bias, running_mean, running_var = f(weight), f(bias), f(running_mean), f(running_var)
    _capi.call("laff_bn_fold", _ptr(weight), _ptr(bias), _ptr(running_mean), _ptr(running_var), float(eps), D,
               _ptr(scale), _ptr(shift), _stream(running_mean))
    return scale, shift


def project(x16: torch.Tensor, w16: torch.Tensor, bias: Optional[torch.Tensor], activation, bn_scale=None, bn_shift=None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = BN(act(x W^T + b)) (TransformNet.forward, model/model.py:257-276). x16 [rows, K], w16 [D, K] 16-bit."""
    x16, w16 = _check_operands(x16, w16)
    rows, K = x16.shape
    D = w16.shape[0]
    if out is None:
        out = torch.empty((rows, D), dtype=torch.float32, device=x16.device)
    act = activation if isinstance(activation, int) else _capi.ACT[activation]
    if rows:
        _capi.call("laff_project", _ptr(x16), _ptr(w16), rows, K, D, x16.stride(0), w16.stride(0), _DT[x16.dtype],
                   _ptr(bias), act, _ptr(bn_scale), _ptr(bn_shift), _ptr(out), out.stride(0), _stream(x16))
    return out


def attention_pool(sources: Sequence[dict], att_weight: torch.Tensor, att_bias: torch.Tensor, heads: int, head_dim: int,
                   with_ave: bool = False, mul: bool = False, omega: float = 1.0, norm_eps: float = 1e-14,
                   out16_dtype=None, want_att: bool = False):
    """LAFF block over L feature sources.

    Each source: {'y': fp32 [rows, D]} (projected) or {'x': fp32 [rows, in_dim], 'bn_scale':..., 'bn_shift':...}
    (no-transform, tiled + BN).  Returns (out fp32 [rows, heads, head_dim], out16 or None, att [rows, heads, L] or None).
    """
    if not 1 <= len(sources) <= _capi.MAX_FEATURES:
        raise LaffError("attention_pool: n_features=%d outside [1, %d]" % (len(sources), _capi.MAX_FEATURES))
    desc = PoolDesc()
    desc.n_features = len(sources)
    desc.heads, desc.head_dim = heads, head_dim
    desc.with_ave, desc.mul = int(bool(with_ave)), int(bool(mul))
    desc.omega = float(omega)
    desc.norm_eps = float(norm_eps)
    att_weight = att_weight.detach().float().contiguous()
    att_bias = att_bias.detach().float().contiguous()
    _need_cuda(att_weight, att_bias)
    desc.att_weight, desc.att_bias = att_weight.data_ptr(), att_bias.data_ptr()
    keep = []
    rows = None
    for l, s in enumerate(sources):
        if "y" in s:
            t = _rowmajor(s["y"])
            desc.src[l].kind, desc.src[l].in_dim = 0, 0
        else:
            t = _rowmajor(s["x"].float() if s["x"].dtype != torch.float32 else s["x"])
            desc.src[l].kind, desc.src[l].in_dim = 1, t.shape[1]
            if s.get("bn_scale") is not None:
                desc.src[l].bn_scale, desc.src[l].bn_shift = s["bn_scale"].data_ptr(), s["bn_shift"].data_ptr()
        _need_cuda(t)
        if t.dtype != torch.float32:
            raise LaffError("pool sources must be fp32")
        keep.append(t)
        desc.src[l].src, desc.src[l].ld = t.data_ptr(), t.stride(0)
        rows = t.shape[0] if rows is None else rows
        if t.shape[0] != rows:
            raise LaffError("pool sources disagree on the number of rows")
    dev = keep[0].device
    D = heads * head_dim
    out = torch.empty((rows, D), dtype=torch.float32, device=dev)
    out16 = None
    o16dt = 0
    if out16_dtype is not None:
        out16_dtype = torch_dtype(out16_dtype)
        out16 = torch.empty((rows, D), dtype=out16_dtype, device=dev)
        o16dt = _DT[out16_dtype]
    att = torch.empty((rows, heads, len(sources)), dtype=torch.float32, device=dev) if want_att else None
    if rows:
        _capi.call("laff_attention_pool", C.byref(desc), rows, _ptr(out), D, _ptr(out16), o16dt, D, _ptr(att),
                   _stream(keep[0]))
    return out.view(rows, heads, head_dim), (None if out16 is None else out16.view(rows, heads, head_dim)), att


def fuse_forward(fc: Sequence[dict], tiled: Sequence[dict], att_weight: torch.Tensor, att_bias: torch.Tensor, heads: int,
                 head_dim: int, want_f32: bool = True, out16_dtype=None, norm_eps: float = 1e-14,
                 out: Optional[torch.Tensor] = None, out16: Optional[torch.Tensor] = None):
    """All projections + LAFF pooling in ONE kernel (laff_fuse_forward; head_dim 512, with_ave = mul = False).

    fc:    [{'x16': [rows, K] 16-bit, 'w16': [D, K] 16-bit, 'bias': fp32 [D] or None, 'activation': name/int,
             'bn_scale': fp32 [D] or None, 'bn_shift': ...}]
    tiled: [{'x': fp32 [rows, in_dim], 'bn_scale': ..., 'bn_shift': ...}]
    out / out16: optional preallocated [rows, heads*head_dim] destinations (row-major; fp32 / 16-bit).
    Returns (out fp32 [rows, heads, head_dim] or None, out16 or None)."""
    if not (1 <= len(fc) <= FUSE_MAX_FC and len(tiled) <= FUSE_MAX_TILED):
        raise LaffError("fuse_forward supports 1..%d projected and up to %d tiled features" % (FUSE_MAX_FC, FUSE_MAX_TILED))
    d = FuseDesc()
    d.n_fc, d.n_tiled, d.heads, d.head_dim = len(fc), len(tiled), heads, head_dim
    d.norm_eps = float(norm_eps)
    att_weight = att_weight.detach().float().contiguous()
    att_bias = att_bias.detach().float().contiguous()
    d.att_weight, d.att_bias = att_weight.data_ptr(), att_bias.data_ptr()
    keep = [att_weight, att_bias]
    rows = None
    dt = None
    for l, f in enumerate(fc):
        x16, w16 = _check_operands(f["x16"], f["w16"])
        keep += [x16, w16]
        dt = x16.dtype if dt is None else dt
        if x16.dtype != dt:
            raise LaffError("all projected features must share one operand dtype")
        rows = x16.shape[0] if rows is None else rows
        e = d.fc[l]
        e.x16, e.ldx, e.w16, e.ldw, e.K = x16.data_ptr(), x16.stride(0), w16.data_ptr(), w16.stride(0), x16.shape[1]
        act = f.get("activation")
        e.activation = act if isinstance(act, int) else _capi.ACT[act]
        for name in ("bias", "bn_scale", "bn_shift"):
            t = f.get(name)
            if t is not None:
                _need_cuda(t)
                keep.append(t)
                setattr(e, name, t.data_ptr())
    for l, f in enumerate(tiled):
        x = _rowmajor(f["x"].float() if f["x"].dtype != torch.float32 else f["x"])
        _need_cuda(x)
        keep.append(x)
        e = d.tiled[l]
        e.x, e.ld, e.in_dim = x.data_ptr(), x.stride(0), x.shape[1]
        if f.get("bn_scale") is not None:
            e.bn_scale, e.bn_shift = f["bn_scale"].data_ptr(), f["bn_shift"].data_ptr()
    d.dtype = _DT[dt]
    dev = keep[2].device
    D = heads * head_dim
    def _dest(t, dtype, what):
        if t.dtype != dtype or t.dim() != 2 or tuple(t.shape) != (rows, D) or t.stride(1) != 1 or not t.is_cuda:
            raise LaffError("fuse_forward: %s must be a CUDA %s [%d, %d] tensor with contiguous rows" % (what, dtype, rows, D))
        return t

    if out is not None:
        out = _dest(out, torch.float32, "out")
    elif want_f32:
        out = torch.empty((rows, D), dtype=torch.float32, device=dev)
    o16 = 0
    if out16 is not None:
        out16 = _dest(out16, out16.dtype if out16_dtype is None else torch_dtype(out16_dtype), "out16")
        o16 = _DT[out16.dtype]
    elif out16_dtype is not None:
        out16_dtype = torch_dtype(out16_dtype)
        out16 = torch.empty((rows, D), dtype=out16_dtype, device=dev)
        o16 = _DT[out16_dtype]
    if out is None and out16 is None:
        raise LaffError("fuse_forward: nothing to write (want_f32 = False and no 16-bit output)")
    if rows:
        _capi.call("laff_fuse_forward", C.byref(d), rows, _ptr(out), 0 if out is None else out.stride(0), _ptr(out16), o16,
                   0 if out16 is None else out16.stride(0), _stream(keep[2]))
    return (None if out is None else out.view(rows, heads, head_dim)), (None if out16 is None else out16.view(rows, heads, head_dim))


def frame_pool(frames: torch.Tensor, att_weight: torch.Tensor, att_bias: float, with_ave: bool = False,
               mul: bool = False, omega: float = 1.0, norm_eps: float = 1e-14) -> torch.Tensor:
    """Frame-level LAFF block: frames fp32 [B, F, dim] -> [B, dim] unit norm."""
    _need_cuda(frames, att_weight)
    frames = frames.float().contiguous()
    B, F, dim = frames.shape
    att_weight = att_weight.detach().float().contiguous().view(-1)
    out = torch.empty((B, dim), dtype=torch.float32, device=frames.device)
    if B:
        _capi.call("laff_frame_pool", _ptr(frames), B, F, dim, _ptr(att_weight), float(att_bias), int(bool(with_ave)),
                   int(bool(mul)), float(omega), float(norm_eps), _ptr(out), out.stride(0), _stream(frames))
    return out


# ----------------------------------------------------------------------------------------------------------------
# text front-end (after tokenisation)
# ----------------------------------------------------------------------------------------------------------------
def _csr(offsets: torch.Tensor, ids: torch.Tensor):
    _need_cuda(offsets, ids)
    return offsets.to(torch.int64).contiguous(), ids.to(torch.int32).contiguous()


class nvtx_range:
    """NVTX range around a stage of the hot path (SURVEY §5 tracing row): shows up in nsys / ncu timelines as
    laff/<name>; a no-op costing well under a microsecond when no profiler is attached.  LAFF_NVTX=0 disables it."""
    enabled = os.environ.get("LAFF_NVTX", "1") != "0"

    def __init__(self, name: str):
        self.name = "laff/" + name

    def __enter__(self):
        if self.enabled and torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if self.enabled and torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


class sm_limit:
    """`with ops.sm_limit(n):` -- the persistent kernels launched inside fill at most n SMs (laff_set_sm_limit); n = 0
    or None leaves the whole device.  Host-side state read at launch time: use from the launching thread only."""

    def __init__(self, sms: Optional[int]):
        self.sms = int(sms or 0)

    def __enter__(self):
        self.prev = _capi.lib().laff_set_sm_limit(self.sms)
        return self

    def __exit__(self, *exc):
        _capi.lib().laff_set_sm_limit(self.prev)
        return False


class SparseRows:
    """A batch of bag-of-words rows kept sparse: CSR token ids (a token that occurs twice is listed twice -- its count)
    instead of the dense [rows, ndims] count matrix of BowVec._encoding (txt2vec.py:56-63; ~8 non-zeros of 3981).

    offsets int64 [rows + 1], ids int32 [n_tokens] (host or device tensors), ndims = vocabulary size; token t of row i
    is ids[offsets[i] - base + ...].  Quacks like the dense tensor where the query path needs it: `.shape`, `.is_cuda`,
    row slicing, `.to(device)`, `.pin_memory()`, `.record_stream()`; `.dense()` expands on the device (the training
    step and the bf16x3 / attention-weight paths still take the dense matrix).  Host-side slices carry only their own
    ids, so a rank (or an H2D chunk) copies 4 bytes per token instead of 4 * ndims bytes per caption."""

    def __init__(self, offsets: torch.Tensor, ids: torch.Tensor, ndims: int, base: int = 0, row_scale: Optional[torch.Tensor] = None):
        self.offsets, self.ids, self.ndims, self.base, self.row_scale = offsets, ids, int(ndims), int(base), row_scale

    @classmethod
    def from_lists(cls, lists, ndims: int) -> "SparseRows":
        off = np.zeros(len(lists) + 1, dtype=np.int64)
        np.cumsum([len(l) for l in lists], out=off[1:])
        ids = np.fromiter((i for l in lists for i in l), dtype=np.int32, count=int(off[-1]))
        return cls(torch.from_numpy(off), torch.from_numpy(ids), ndims)

    @classmethod
    def from_dense(cls, counts: torch.Tensor) -> "SparseRows":
        """Exact CSR form of a dense count matrix with non-negative integer entries (host tensor)."""
        c = counts.detach().cpu()
        if not bool(((c >= 0) & (c == c.round())).all()):
            raise LaffError("SparseRows.from_dense: a BoW count matrix has non-negative integer entries")
        r, col = torch.nonzero(c, as_tuple=True)
        rep = c[r, col].to(torch.int64)
        ids = torch.repeat_interleave(col, rep).to(torch.int32)
        per_row = torch.zeros(c.shape[0], dtype=torch.int64).index_add_(0, r, rep)
        off = torch.zeros(c.shape[0] + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(per_row, 0)
        return cls(off, ids, c.shape[1])

    @property
    def shape(self):
        return (self.offsets.numel() - 1, self.ndims)

    @property
    def is_cuda(self):
        return self.offsets.is_cuda

    @property
    def device(self):
        return self.offsets.device

    def nbytes(self) -> int:
        return self.offsets.numel() * 8 + self.ids.numel() * 4 + (0 if self.row_scale is None else self.row_scale.numel() * 4)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key) -> "SparseRows":
        if not isinstance(key, slice) or key.step not in (None, 1):
            raise LaffError("SparseRows supports contiguous row slices only")
        lo, hi, _ = key.indices(self.shape[0])
        hi = max(hi, lo)
        off = self.offsets[lo:hi + 1]
        rs = None if self.row_scale is None else self.row_scale[lo:hi]
        if self.offsets.is_cuda:                       # no host read: keep the id array whole
            return SparseRows(off, self.ids, self.ndims, self.base, rs)
        a, b = int(off[0]) - self.base, int(off[-1]) - self.base
        return SparseRows(off, self.ids[a:b], self.ndims, self.base + a, rs)

    def to(self, device, non_blocking: bool = False) -> "SparseRows":
        return SparseRows(self.offsets.to(device, non_blocking=non_blocking), self.ids.to(device, non_blocking=non_blocking),
                          self.ndims, self.base, None if self.row_scale is None else self.row_scale.to(device, non_blocking=non_blocking))

    def pin_memory(self) -> "SparseRows":
        return SparseRows(self.offsets.contiguous().pin_memory(), self.ids.contiguous().pin_memory(), self.ndims, self.base,
                          None if self.row_scale is None else self.row_scale.pin_memory())

    def float(self) -> "SparseRows":
        return self

    def record_stream(self, stream) -> None:
        for t in (self.offsets, self.ids, self.row_scale):
            if t is not None and t.is_cuda:
                t.record_stream(stream)

    def dense(self) -> torch.Tensor:
        _need_cuda(self.offsets)
        off = (self.offsets - self.base) if self.base else self.offsets
        out = bow_counts(off, self.ids, self.ndims)
        return out if self.row_scale is None else out * self.row_scale[:, None]


def bow_project(x: "SparseRows", wt: torch.Tensor, bias: Optional[torch.Tensor], activation, bn_scale=None, bn_shift=None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = BN(act(counts @ W^T + b)) of a sparse BoW batch as a gather-sum over the rows of wt = W^T (fp32 [ndims, D])
    (laff_bow_project; the reference's BoWTxtEncoder + TransformNet, model/model.py:399-417, :257-276)."""
    _need_cuda(x.offsets, x.ids, wt)
    if wt.dtype != torch.float32 or wt.dim() != 2 or wt.stride(1) != 1 or wt.shape[0] != x.ndims:
        raise LaffError("bow_project: wt must be fp32 [ndims = %d, D] row-major, got %s" % (x.ndims, tuple(wt.shape)))
    rows, D = x.shape[0], wt.shape[1]
    offsets, ids = _csr(x.offsets, x.ids)
    if out is None:
        out = torch.empty((rows, D), dtype=torch.float32, device=wt.device)
    act = activation if isinstance(activation, int) else _capi.ACT[activation]
    _capi.call("laff_bow_project", _ptr(offsets), _ptr(ids), x.base, rows, x.ndims, _ptr(wt), wt.stride(0), D, _ptr(bias), act,
               _ptr(bn_scale), _ptr(bn_shift), _ptr(x.row_scale), _ptr(out), out.stride(0), _stream(wt))
    return out


def bow_counts(offsets: torch.Tensor, ids: torch.Tensor, ndims: int) -> torch.Tensor:
    """BowVec._encoding for a batch: CSR token ids -> fp32 [rows, ndims] count vectors (laff_bow_counts)."""
    offsets, ids = _csr(offsets, ids)
    rows = offsets.numel() - 1
    out = torch.empty((rows, ndims), dtype=torch.float32, device=offsets.device)
    _capi.call("laff_bow_counts", _ptr(offsets), _ptr(ids), rows, int(ndims), _ptr(out), out.stride(0), _stream(offsets))
    return out


def gather_mean(table: torch.Tensor, offsets: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """W2Vec._encoding for a batch: mean (fp64 accumulation, list order) of table rows per caption (laff_gather_mean)."""
    _need_cuda(table)
    table = _rowmajor(table.float() if table.dtype != torch.float32 else table)
    offsets, ids = _csr(offsets, ids)
    rows = offsets.numel() - 1
    out = torch.empty((rows, table.shape[1]), dtype=torch.float32, device=table.device)
    _capi.call("laff_gather_mean", _ptr(table), table.stride(0), table.shape[0], _ptr(offsets), _ptr(ids), rows, table.shape[1],
               _ptr(out), out.stride(0), _stream(table))
    return out


def gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """nn.Embedding lookup: fp32 [n, dim] = table[ids] (laff_gather_rows)."""
    _need_cuda(table, ids)
    table = _rowmajor(table.float() if table.dtype != torch.float32 else table)
    ids = ids.to(torch.int32).contiguous().view(-1)
    out = torch.empty((ids.numel(), table.shape[1]), dtype=torch.float32, device=table.device)
    _capi.call("laff_gather_rows", _ptr(table), table.stride(0), table.shape[0], _ptr(ids), ids.numel(), table.shape[1],
               _ptr(out), out.stride(0), _stream(table))
    return out


def gru_cell(gi: torch.Tensor, gh: torch.Tensor, h_prev: torch.Tensor, lengths: torch.Tensor, t: int, h_out: torch.Tensor,
             sum_out: Optional[torch.Tensor] = None, last_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One nn.GRU step over a batch of packed sequences (laff_gru_cell).  gi: fp32 [B, 3H] view (row pitch free) of the
    input-side pre-activations of step t; gh: fp32 [B, 3H] hidden-side pre-activations."""
    _need_cuda(gi, gh, h_prev, lengths, h_out, sum_out, last_out)
    B, H = h_prev.shape
    if gi.shape != (B, 3 * H) or gh.shape != (B, 3 * H) or gi.stride(1) != 1 or gh.stride(1) != 1:
        raise LaffError("gru_cell: gi / gh must be [B, 3H] with contiguous rows")
    _capi.call("laff_gru_cell", _ptr(gi), gi.stride(0), _ptr(gh), gh.stride(0), _ptr(h_prev), _ptr(lengths), int(t), B, H,
               _ptr(h_out), _ptr(sum_out), _ptr(last_out), _stream(h_prev))
    return h_out


def mean_over_length(x: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, lengths)
    _capi.call("laff_mean_over_length", _ptr(x), _ptr(lengths), x.shape[0], x.shape[1], _stream(x))
    return x


# ----------------------------------------------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------------------------------------------
def mrl_forward_backward(txt: torch.Tensor, vis: torch.Tensor, margin: float, max_violation: bool, direction: str,
                         cost_style: str, need_grad: bool = True):
    """Sum over heads of MarginRankingLoss(s=txt[:,h], im=vis[:,h]); returns (loss scalar tensor, d_txt, d_vis)."""
    _need_cuda(txt, vis)
    if txt.dim() == 2:
        txt, vis = txt.unsqueeze(1), vis.unsqueeze(1)
    txt = txt.detach().float().contiguous()
    vis = vis.detach().float().contiguous()
    B, H, dh = txt.shape
    loss = torch.empty((), dtype=torch.float32, device=txt.device)
    d_txt = torch.empty_like(txt) if need_grad else None
    d_vis = torch.empty_like(vis) if need_grad else None
    nbytes = _capi.lib().laff_mrl_workspace_bytes(B, H, dh)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=txt.device)
    _capi.call("laff_mrl_forward_backward", _ptr(txt), _ptr(vis), B, H, dh, float(margin), int(bool(max_violation)),
               _capi.DIRECTION[direction], int(cost_style != "sum"), _ptr(loss), _ptr(d_txt), _ptr(d_vis), _ptr(ws),
               nbytes, _stream(txt))
    return loss, d_txt, d_vis


def dsl_forward_backward(txt: torch.Tensor, vis: torch.Tensor, temp: float = 1000.0, need_grad: bool = True):
    """Sum over heads of DualSoftmaxLoss(s=txt[:,h], im=vis[:,h], temp) (loss.py:291-310); (loss, d_txt, d_vis)."""
    _need_cuda(txt, vis)
    if txt.dim() == 2:
        txt, vis = txt.unsqueeze(1), vis.unsqueeze(1)
    txt = txt.detach().float().contiguous()
    vis = vis.detach().float().contiguous()
    B, H, dh = txt.shape
    loss = torch.empty((), dtype=torch.float32, device=txt.device)
    d_txt = torch.empty_like(txt) if need_grad else None
    d_vis = torch.empty_like(vis) if need_grad else None
    nbytes = _capi.lib().laff_mrl_workspace_bytes(B, H, dh)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=txt.device)
    _capi.call("laff_dsl_forward_backward", _ptr(txt), _ptr(vis), B, H, dh, float(temp), _ptr(loss), _ptr(d_txt), _ptr(d_vis),
               _ptr(ws), nbytes, _stream(txt))
    return loss, d_txt, d_vis


def mrl_score_forward_backward(score: torch.Tensor, margin: float, max_violation: bool, direction: str,
                               cost_style: str, need_grad: bool = True):
    _need_cuda(score)
    # the C entry point uses one pitch for the scores and their gradient: a strided view of a wider matrix would make
    # it write past the contiguous [B, B] gradient, so the scores are made contiguous first
    score = score.detach().float().contiguous()
    B = score.shape[0]
    loss = torch.empty((), dtype=torch.float32, device=score.device)
    d_score = torch.empty((B, B), dtype=torch.float32, device=score.device) if need_grad else None
    _capi.call("laff_mrl_score_forward_backward", _ptr(score), B, score.stride(0), float(margin),
               int(bool(max_violation)), _capi.DIRECTION[direction], int(cost_style != "sum"), _ptr(loss),
               _ptr(d_score), _stream(score))
    return loss, d_score


# ----------------------------------------------------------------------------------------------------------------
# training step (T1 / N4)
# ----------------------------------------------------------------------------------------------------------------
def transform_train_forward(src: torch.Tensor, D: int, p_drop: float, seed: int, bn=None, momentum: float = 0.1,
                            seed_dev: Optional[torch.Tensor] = None):
    """Dropout + train-mode BatchNorm1d after the activation (laff_transform_train_forward).
    src fp32 [B, D] (activated projection) or [B, in_dim] (raw feature tiled D / in_dim times).
    bn: None or an nn.BatchNorm1d whose running statistics are updated in place.
    Returns (y [B, D], mask uint8 [B, D] or None, save_mean or None, save_invstd or None)."""
    _need_cuda(src)
    src = _rowmajor(src.float() if src.dtype != torch.float32 else src)
    B = src.shape[0]
    dev = src.device
    y = torch.empty((B, D), dtype=torch.float32, device=dev)
    mask = torch.empty((B, D), dtype=torch.uint8, device=dev) if p_drop > 0 else None
    sm = si = None
    gamma = beta = rm = rv = None
    eps = 1e-5
    if bn is not None:
        sm = torch.empty(D, dtype=torch.float32, device=dev)
        si = torch.empty(D, dtype=torch.float32, device=dev)
        gamma, beta, rm, rv, eps = bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps
        momentum = bn.momentum if bn.momentum is not None else momentum
    _capi.call("laff_transform_train_forward", _ptr(src), src.stride(0), src.shape[1], B, D, float(p_drop),
               int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(seed_dev), _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), float(momentum), float(eps),
               int(bn is not None), _ptr(y), y.stride(0), _ptr(mask), _ptr(sm), _ptr(si), _stream(src))
    return y, mask, sm, si


def transform_train_backward(dy: torch.Tensor, a: Optional[torch.Tensor], tiled_x: Optional[torch.Tensor], mask, p_drop: float,
                             activation, bn, save_mean, save_invstd, want_dz: bool = True, dgamma=None, dbeta=None, dbias=None):
    """Backward of transform_train_forward down to the GEMM output (laff_transform_train_backward).
    Returns dz [B, D] (None for a tiled feature); writes dgamma / dbeta / dbias [D] when given."""
    _need_cuda(dy, a, tiled_x)
    B, D = dy.shape
    dz = torch.empty((B, D), dtype=torch.float32, device=dy.device) if want_dz else None
    act = activation if isinstance(activation, int) else _capi.ACT[activation]
    _capi.call("laff_transform_train_backward", _ptr(dy), dy.stride(0), _ptr(a), 0 if a is None else a.stride(0), _ptr(tiled_x),
               0 if tiled_x is None else tiled_x.stride(0), 0 if tiled_x is None else tiled_x.shape[1], _ptr(mask), float(p_drop),
               act, int(bn is not None), _ptr(bn.weight if bn is not None else None), _ptr(save_mean), _ptr(save_invstd), B, D,
               _ptr(dz), 0 if dz is None else dz.stride(0), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _stream(dy))
    return dz


def attention_pool_backward(ys: Sequence[torch.Tensor], att_weight: torch.Tensor, att_bias: torch.Tensor, heads: int, head_dim: int,
                            dout: torch.Tensor, dw: torch.Tensor, dc: torch.Tensor, norm_eps: float = 1e-14, with_ave: bool = False,
                            mul: bool = False, omega: float = 0.0):
    """Backward of the LAFF block.  ys: the L inputs fp32 [rows, H*d_h]; dout [rows, H*d_h];
    dw [H, d_h] / dc [H] receive the gradients of the logit weights / biases.  Returns the list of dy_l."""
    _need_cuda(dout, att_weight, att_bias, dw, dc, *ys)
    rows, D = dout.shape
    dev = dout.device
    ys = [_rowmajor(y) for y in ys]
    dys = [torch.empty((rows, D), dtype=torch.float32, device=dev) for _ in ys]
    n = len(ys)
    yp = (C.c_void_p * n)(*[y.data_ptr() for y in ys])          # host arrays: the entry point copies them into the
    dp = (C.c_void_p * n)(*[d.data_ptr() for d in dys])         # kernel's by-value argument block
    lds = (C.c_longlong * n)(*[y.stride(0) for y in ys])
    dw_part = torch.empty(rows * D, dtype=torch.float32, device=dev)
    dc_part = torch.empty(rows * heads, dtype=torch.float32, device=dev)
    dout = _rowmajor(dout)
    _capi.call("laff_attention_pool_backward", yp, lds, n, heads, head_dim, _ptr(att_weight), _ptr(att_bias),
               _ptr(dout), dout.stride(0), rows, float(norm_eps), int(bool(with_ave)), int(bool(mul)), float(omega), dp,
               _ptr(dw_part), _ptr(dc_part), _ptr(dw), _ptr(dc), _stream(dout))
    return dys


def transpose_16(x: torch.Tensor, dtype=torch.bfloat16, terms: int = 1, side: int = 0) -> torch.Tensor:
    """fp32 [R, C] -> 16-bit [C, pad8(R) * terms] (laff_transpose_16): K-major operand of a product contracted over R."""
    _need_cuda(x)
    x = _rowmajor(x.float() if x.dtype != torch.float32 else x)
    R, Cn = x.shape
    kpad = (R + 7) // 8 * 8
    dtype = torch_dtype(dtype)
    out = torch.empty((Cn, kpad * terms), dtype=dtype, device=x.device)
    _capi.call("laff_transpose_16", _ptr(x), x.stride(0), R, Cn, _DT[dtype], int(terms), int(side), _ptr(out), out.stride(0),
               _stream(x))
    return out


def fold_tiles(dz: torch.Tensor, in_dim: int) -> torch.Tensor:
    """Gradient w.r.t. a feature that was tiled over the heads: [B, D] -> [B, in_dim] (laff_fold_tiles)."""
    _need_cuda(dz)
    B, D = dz.shape
    dx = torch.empty((B, in_dim), dtype=torch.float32, device=dz.device)
    _capi.call("laff_fold_tiles", _ptr(dz), dz.stride(0), B, D, int(in_dim), _ptr(dx), dx.stride(0), _stream(dz))
    return dx


def frame_pool_backward(frames: torch.Tensor, att_weight: torch.Tensor, dout: torch.Tensor, dw: torch.Tensor, dc: torch.Tensor,
                        norm_eps: float = 1e-14) -> None:
    """Backward of the frame-level LAFF block w.r.t. its logit weight / bias (laff_frame_pool_backward)."""
    _need_cuda(frames, att_weight, dout, dw, dc)
    frames = frames.float().contiguous()
    B, F, dim = frames.shape
    att_weight = att_weight.detach().float().contiguous().view(-1)
    dout = _rowmajor(dout)
    dw_part = torch.empty(B * dim, dtype=torch.float32, device=frames.device)
    dc_part = torch.empty(B, dtype=torch.float32, device=frames.device)
    _capi.call("laff_frame_pool_backward", _ptr(frames), B, F, dim, _ptr(att_weight), _ptr(dout), dout.stride(0), float(norm_eps),
               _ptr(dw_part), _ptr(dc_part), _ptr(dw), _ptr(dc), _stream(frames))


def gru_cell_backward(gi, gh, h_prev, dmean, dlast, dh_gemm, lengths, t: int, dh_carry, dgi, dgh) -> None:
    """One BPTT step of the GRU (laff_gru_cell_backward); gi / gh / dgi / dgh are [B, 3H] views with contiguous rows."""
    _need_cuda(gi, gh, h_prev, dmean, dlast, dh_gemm, lengths, dh_carry, dgi, dgh)
    B, H = h_prev.shape
    _capi.call("laff_gru_cell_backward", _ptr(gi), gi.stride(0), _ptr(gh), gh.stride(0), _ptr(h_prev), _ptr(dmean), _ptr(dlast),
               _ptr(dh_gemm), _ptr(lengths), int(t), B, H, _ptr(dh_carry), _ptr(dgi), dgi.stride(0), _ptr(dgh), dgh.stride(0),
               _stream(h_prev))


def scatter_add_rows(dx: torch.Tensor, ids: torch.Tensor, table_grad: torch.Tensor) -> None:
    """nn.Embedding backward: table_grad[ids[i]] += dx[i] (laff_scatter_add_rows); table_grad must be zeroed by the caller."""
    _need_cuda(dx, ids, table_grad)
    dx = _rowmajor(dx)
    ids = ids.to(torch.int32).contiguous().view(-1)
    _capi.call("laff_scatter_add_rows", _ptr(dx), dx.stride(0), _ptr(ids), ids.numel(), dx.shape[1], table_grad.shape[0],
               _ptr(table_grad), table_grad.stride(0), _stream(dx))


def column_sum(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[c] = sum_r x[r, c], fixed order (laff_column_sum)."""
    _need_cuda(x, out)
    x = _rowmajor(x)
    if out is None:
        out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    _capi.call("laff_column_sum", _ptr(x), x.stride(0), x.shape[0], x.shape[1], _ptr(out), _stream(x))
    return out
