"""Drop-in for the reference's ``loss.py`` hot-path surface (loss.py:8-13, :30-34, :68-135, :138-200).

Same names, argument meaning and error behaviour as the reference; the arithmetic runs in the sm_100a kernels behind
``include/laff_b200.h``.  CUDA tensors only (the reference moves tensors to its global ``device`` itself; here the
caller's tensors decide the device), no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

# Operand precision of the tensor-core similarity GEMM:
#   'fp16'   fp16 operands, fp32 accumulate (default).  Same tensor-core rate as bf16 (`kind::f16` covers both) with an
#            11-bit instead of an 8-bit significand: on the reference-trained fixture it moves 0.1-0.7 % of the ranks
#            against the fp32 reference where bf16 moves 1.8-4.4 % (tests/test_gpu_trained.py,
#            profiles/r02_parity_matrix.jsonl).  Unit-norm embeddings always fit fp16's range; raw features are
#            saturated at +-65504 by the cast.
#   'bf16'   bf16 operands (what BASELINE.json's north_star names; kept selectable)
#   'bf16x3' 3-term split (hi*hi + lo*hi + hi*lo): near-fp32 products at 3x the MMA work -- ranks identical to the
#            fp32 reference on the trained fixture; for small problems and as the parity reference of bench.py
_PRECISION = "fp16"


def set_precision(p: str) -> None:
    global _PRECISION
    if p not in ("bf16", "fp16", "bf16x3"):
        raise ValueError("precision must be 'bf16', 'fp16' or 'bf16x3'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def operand_dtype(precision: str = None) -> torch.dtype:
    """The 16-bit tensor type in which embeddings of the given (default: current) precision are kept."""
    return torch.float16 if (precision or _PRECISION) == "fp16" else torch.bfloat16


def l2norm(X: torch.Tensor, eps: float = 1e-13, dim: int = 1) -> torch.Tensor:
    """L2-normalise along ``dim``: X / (||X|| + eps + 1e-14)  (loss.py:8-13)."""
    if dim < 0:
        dim += X.dim()
    if dim != X.dim() - 1:
        return l2norm(X.transpose(dim, -1), eps, -1).transpose(dim, -1)
    shape = X.shape
    X2 = X.reshape(-1, shape[-1])
    out = ops.l2norm_quantize(X2, 1, torch.float32, eps=eps + 1e-14)
    return out.reshape(shape)


def _operands(query: torch.Tensor, retrio: torch.Tensor, heads: int, precision: str, normalise: bool = True):
    """l2norm per head (loss.py:32) and round to the tensor-core operand type."""
    if precision == "bf16x3":
        qn = ops.l2norm_quantize(query, heads, torch.float32, normalise=normalise)
        rn = ops.l2norm_quantize(retrio, heads, torch.float32, normalise=normalise)
        return ops.split3_16(qn, 0, torch.bfloat16), ops.split3_16(rn, 1, torch.bfloat16)
    dt = torch.bfloat16 if precision == "bf16" else torch.float16
    return (ops.l2norm_quantize(query, heads, dt, normalise=normalise),
            ops.l2norm_quantize(retrio, heads, dt, normalise=normalise))


def cosine_sim(query: torch.Tensor, retrio: torch.Tensor) -> torch.Tensor:
    """Cosine similarity between all query / retrio pairs (loss.py:30-34): l2norm both, query.mm(retrio.t())."""
    q16, r16 = _operands(query, retrio, 1, _PRECISION)
    return ops.sim_dense(q16, r16, 1.0)


class _MRLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, im, margin, max_violation, direction, cost_style):
        need = s.requires_grad or im.requires_grad
        loss, d_s, d_im = ops.mrl_forward_backward(s, im, margin, max_violation, direction, cost_style, need_grad=need)
        ctx.shape = s.shape
        ctx.save_for_backward(d_s, d_im) if need else None
        return loss

    @staticmethod
    def backward(ctx, g):
        d_s, d_im = ctx.saved_tensors
        return g * d_s.reshape(ctx.shape), g * d_im.reshape(ctx.shape), None, None, None, None


class _MRLScoreFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, margin, max_violation, direction, cost_style):
        loss, d = ops.mrl_score_forward_backward(score, margin, max_violation, direction, cost_style,
                                                 need_grad=score.requires_grad)
        if score.requires_grad:
            ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return g * d, None, None, None, None


class _DSLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, im, temp):
        need = s.requires_grad or im.requires_grad
        loss, d_s, d_im = ops.dsl_forward_backward(s, im, temp, need_grad=need)
        ctx.shape = s.shape
        ctx.save_for_backward(d_s, d_im) if need else None
        return loss

    @staticmethod
    def backward(ctx, g):
        d_s, d_im = ctx.saved_tensors
        return g * d_s.reshape(ctx.shape), g * d_im.reshape(ctx.shape), None


class DualSoftmaxLoss(nn.Module):
    """loss.py:291-310.  forward(s, im, temp=1000) on [B, d] embeddings; [B, H, d] inputs return the sum over heads
    (model/model.py:2036-2038 applies the criterion head by head)."""

    def forward(self, s, im, temp=1000):
        return _DSLFunction.apply(s, im, float(temp))


class MarginRankingLoss(nn.Module):
    """Margin ranking loss on (sentence, image/video) embedding batches (loss.py:68-135).

    ``forward(s, im)`` accepts [B, d] like the reference, or [B, H, d] multi-space embeddings, in which case it
    returns the sum over heads that W2VVPP.compute_loss builds with a Python loop (model/model.py:857-858).
    """

    def __init__(self, margin=0, measure="cosine", max_violation=False, cost_style="sum", direction="bidir",
                 device=torch.device("cpu")):
        super().__init__()
        self.margin = margin
        self.cost_style = cost_style
        self.direction = direction
        if measure == "cosine":
            self.sim = cosine_sim
        elif measure == "hist":
            raise Exception("measure 'hist' is outside the LAFF hot path (configs use 'cosine', base_config.py:92)")
        else:
            raise Exception("Not implemented.")
        self.max_violation = max_violation

    def forward(self, s, im):
        if self.direction not in ("i2t", "t2i", "bidir"):
            # the reference computes nothing for an unknown direction and returns 0 (loss.py:127-130)
            return torch.zeros((), device=s.device)
        return _MRLFunction.apply(s, im, float(self.margin), bool(self.max_violation), self.direction, self.cost_style)


class MarginRankingLossWithScore(nn.Module):
    """The same hinge on a given score matrix (loss.py:138-200)."""

    def __init__(self, margin=0, max_violation=False, cost_style="sum", direction="bidir", device=torch.device("cpu")):
        super().__init__()
        self.margin = margin
        self.cost_style = cost_style
        self.direction = direction
        self.max_violation = max_violation
        self.device = device

    def forward(self, score):
        if self.direction not in ("i2t", "t2i", "bidir"):
            return torch.zeros((), device=score.device)
        return _MRLScoreFunction.apply(score, float(self.margin), bool(self.max_violation), self.direction,
                                       self.cost_style)
