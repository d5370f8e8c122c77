"""Drop-in for the reference's ``evaluation.py`` (evaluation.py:11-16, :44-61, :64-89, :92-109).

Inputs may be numpy arrays (as the reference's callers pass) or CUDA tensors; numpy inputs are copied to the current
CUDA device, the work runs in the sm_100a kernels, and the same Python types as the reference come back (ndarray /
tuple of floats).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi, ops
from . import loss as _loss


def _dev() -> torch.device:
    if not torch.cuda.is_available():
        raise _capi.LaffError("laff_b200.evaluation needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_cuda(x, dtype=torch.float32) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=_dev() if not x.is_cuda else x.device, dtype=dtype)
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(_dev())


def l2norm(X):
    """X / (||X|| + 1e-10), rows (evaluation.py:11-16)."""
    out = ops.l2norm_quantize(_to_cuda(X), 1, torch.float32, eps=1e-10)
    return out if isinstance(X, torch.Tensor) else out.cpu().numpy()


def cosine_sim(query_embs, retro_embs):
    """evaluation.py:44-50: l2norm (eps 1e-10) both sides then the dot products."""
    q = _to_cuda(query_embs)
    r = _to_cuda(retro_embs)
    prec = _loss.get_precision()
    if prec == "bf16x3":
        qn = ops.l2norm_quantize(q, 1, torch.float32, eps=1e-10)
        rn = ops.l2norm_quantize(r, 1, torch.float32, eps=1e-10)
        q16, r16 = ops.split3_16(qn, 0), ops.split3_16(rn, 1)
    else:
        dt = torch.bfloat16 if prec == "bf16" else torch.float16
        q16 = ops.l2norm_quantize(q, 1, dt, eps=1e-10)
        r16 = ops.l2norm_quantize(r, 1, dt, eps=1e-10)
    out = ops.sim_dense(q16, r16, 1.0)
    both_t = isinstance(query_embs, torch.Tensor) and isinstance(retro_embs, torch.Tensor)
    return out if both_t else out.cpu().numpy()


def compute_sim(query_embs, retro_embs, measure="cosine", device=torch.device("cpu")):
    """evaluation.py:53-61 (same error behaviour)."""
    if measure == "cosine":
        return cosine_sim(query_embs, retro_embs)
    elif measure == "hist":
        raise Exception("measure 'hist' is outside the LAFF hot path")
    elif measure == "euclidean":
        raise Exception("Not implemented")
    else:
        raise Exception("%s is invalid" % measure)


def metrics_from_rank0(rank0) -> tuple:
    """(r1, r5, r10, medr, meanr, mir) from 0-based ranks, computed on device (evaluation.py:81-89)."""
    r = _to_cuda(rank0, torch.int32)
    m = ops.rank_metrics(r).cpu().tolist()
    return (m[0], m[1], m[2], m[3], m[4], m[5])


def eval_qry2retro(qry2retro_sim, n_qry=1):
    """Query -> retrieval metrics of a (n_qry*N, N) similarity matrix (evaluation.py:64-89); ground truth of row i is
    column i / n_qry.  Exact ties are ordered by (score desc, index desc), see DESIGN.md."""
    s = _to_cuda(qry2retro_sim)
    assert s.shape[0] / s.shape[1] == n_qry, tuple(s.shape)
    gt = (torch.arange(s.shape[0], device=s.device) // n_qry).to(torch.int32)
    rank0, _, _ = ops.rank_from_scores(s, gt, 0)
    return metrics_from_rank0(rank0)


def eval(label_matrix):
    """(r1, r5, r10, medr, meanr, mir, mAP) of a 0/1 label matrix in ranked order (evaluation.py:92-109)."""
    if isinstance(label_matrix, torch.Tensor):
        lab = (label_matrix.to(_dev() if not label_matrix.is_cuda else label_matrix.device) == 1).to(torch.uint8)
    else:
        lab = torch.from_numpy(np.ascontiguousarray(label_matrix).astype(int) == 1).to(torch.uint8).to(_dev())
    lab = lab.contiguous()
    Q, V = lab.shape
    rank0 = torch.empty(Q, dtype=torch.int32, device=lab.device)
    ap = torch.empty(Q, dtype=torch.float64, device=lab.device)
    out = torch.empty(8, dtype=torch.float64, device=lab.device)
    st = C.c_void_p(torch.cuda.current_stream(lab.device).cuda_stream)
    _capi.call("laff_label_metrics", C.c_void_p(lab.data_ptr()), Q, V, lab.stride(0), C.c_void_p(rank0.data_ptr()),
               C.c_void_p(ap.data_ptr()), C.c_void_p(out.data_ptr()), st)
    if bool((rank0 < 0).any()):
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # a row without ground truth (evaluation.py:99)
    m = out.cpu().tolist()
    # evaluation.eval works on 1-based ranks: medr = floor(median(rank1)) = floor(median(rank0)) + 1, meanr likewise
    return (m[0], m[1], m[2], m[3], m[4], m[5], m[6])
