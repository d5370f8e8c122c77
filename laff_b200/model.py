"""Drop-in modules for the reference's model-side hot path (model/model.py, model/Attention.py).

Class names, constructor arguments, ``state_dict`` keys/shapes and forward signatures follow the reference so that
``predictor.py`` / ``trainer.validate`` can use these classes with reference checkpoints
(``load_state_dict(checkpoint['model'], strict=False)``, predictor.py:167); the arithmetic runs in the sm_100a kernels
behind ``include/laff_b200.h``:

  TransformNet                        model/model.py:211-276
  Attention_1                         model/Attention.py:40-105
  Multi_head_MyApply_Attention        model/Attention.py:473-552
  VisMutiTransformNet                 model/model.py:1787-1827
  VisMutiTransformNetAddAttnetion     model/model.py:1830-1881
  MultiScaleTxtEncoderAttention       model/model.py:1641-1709   (text features are inputs at this tier)
  VisMutiTransformNetPlusFrameFeat    model/model.py:2101-2194
  W2VVPP_MultiHeadAttention ('LAFF'), W2VVPP_MutiVisFrameFeat ('FrameLAFF'), get_model   model/model.py:2501-2519

Scope: inference (``eval()``) forward, similarity, ranking, and the margin-ranking loss with its gradient w.r.t. the
embeddings.  Train-mode forward of the fusion nets (dropout, batch-statistics BN, backward through the projections)
is the "next" row N4 of SURVEY §8(f) and raises NotImplementedError instead of silently running another path.
"""
from __future__ import annotations

import contextlib

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from . import loss as _loss
from .loss import DualSoftmaxLoss, MarginRankingLoss, MarginRankingLossWithScore

_PARAM_EPOCH = 0  # bumped by the device optimizer step, which updates parameters through raw pointers

_ROW_CHUNK = 32768  # rows fused per pass: bounds the projected-feature scratch to L * 512 MB
_FUSED_ROW_CHUNK = 131072  # rows per laff_fuse_forward launch: bounds the 16-bit input copies (1.3 GB for the video net)


def _cuda_device(hint: Optional[torch.device] = None) -> torch.device:
    if hint is not None and hint.type == "cuda":
        return hint
    if not torch.cuda.is_available():
        raise ops.LaffError("laff_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _op_dtype(precision: str) -> torch.dtype:
    return torch.float16 if precision == "fp16" else torch.bfloat16


def _initialize_weights(m):
    """model/model.py:51-60."""
    if type(m) == nn.Linear:
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif type(m) == nn.BatchNorm1d:
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)


class TransformNet(nn.Module):
    """fc_layers = (dim_in, dim_out): FC -> activation -> dropout -> BatchNorm, each optional (model/model.py:211-276)."""

    def __init__(self, fc_layers, opt=None, dropout=None, batch_norm=None, activation=None, fc=True):
        super().__init__()
        if opt is not None:
            if batch_norm is None:
                batch_norm = opt.batch_norm
            if activation is None:
                activation = opt.activation
            if dropout is None:
                dropout = opt.dropout
        self.fc1 = nn.Linear(fc_layers[0], fc_layers[1]) if fc else None
        self.bn1 = nn.BatchNorm1d(fc_layers[1]) if batch_norm else None
        self.activation_name = activation if activation in ("tanh", "relu", "sigmoid") else None
        self.dropout_p = dropout if (dropout is not None and dropout > 1e-3) else None
        self.out_dim = fc_layers[1]
        self._cache = {}
        self.apply(_initialize_weights)

    # -- prepared (16-bit / folded) parameters, rebuilt when the parameters change -----------------------------
    def _version(self):
        return (_PARAM_EPOCH,) + tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def prepared(self, precision: str):
        key = (precision, self._version())
        if self._cache.get("key") != key:
            c = {"key": key}
            if self.fc1 is not None:
                w = self.fc1.weight.detach()
                if precision == "bf16x3":
                    c["w16"] = ops.split3_16(w, 1, torch.bfloat16)
                else:
                    c["w16"] = ops.cast_pad_16(w, _op_dtype(precision))
                c["bias"] = self.fc1.bias.detach().float().contiguous()
            if self.bn1 is not None:
                c["bn_scale"], c["bn_shift"] = ops.bn_fold(self.bn1.weight, self.bn1.bias, self.bn1.running_mean,
                                                           self.bn1.running_var, self.bn1.eps)
            if "sparse" in self._cache:
                c["sparse"] = self._cache["sparse"]
            self._cache = c
        return self._cache

    def prepared_sparse(self):
        """W^T as fp32 [d_in, D] for the gather-sum projection of a sparse BoW feature (rebuilt when W changes)."""
        key = ("sparse", self._version())
        c = self._cache.get("sparse")
        if c is None or c["key"] != key:
            c = {"key": key, "wt": self.fc1.weight.detach().float().t().contiguous(),
                 "bias": self.fc1.bias.detach().float().contiguous()}
            if self.bn1 is not None:
                c["bn_scale"], c["bn_shift"] = ops.bn_fold(self.bn1.weight, self.bn1.bias, self.bn1.running_mean,
                                                           self.bn1.running_var, self.bn1.eps)
            self._cache["sparse"] = c
        return c

    def project_sparse(self, x: "ops.SparseRows", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """y = BN(act(counts W^T + b)) of a CSR BoW batch without the dense count matrix (ops.bow_project)."""
        c = self.prepared_sparse()
        return ops.bow_project(x, c["wt"], c["bias"], self.activation_name, c.get("bn_scale"), c.get("bn_shift"), out=out)

    def project(self, x: torch.Tensor, precision: str, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """y = BN(act(x W^T + b)) for an fc feature; x fp32 [rows, d_in] on the device."""
        if isinstance(x, ops.SparseRows):
            return self.project_sparse(x, out=out)
        c = self.prepared(precision)
        if precision == "bf16x3":
            x16 = ops.split3_16(x, 0, torch.bfloat16)
        else:
            x16 = ops.cast_pad_16(x, _op_dtype(precision))
        return ops.project(x16, c["w16"], c["bias"], self.activation_name, c.get("bn_scale"), c.get("bn_shift"), out=out)

    def forward(self, input_x):
        if self.training and (self.dropout_p is not None or self.bn1 is not None):
            raise NotImplementedError("laff_b200.TransformNet: train-mode forward (dropout / batch-stat BN) is not part "
                                      "of the hot path built so far (SURVEY §8f N4); call .eval()")
        x = input_x.to(_cuda_device(self._param_device()), non_blocking=True).float()
        precision = _loss.get_precision()
        if self.fc1 is not None:
            return self.project(x, precision)
        if self.bn1 is not None:
            c = self.prepared(precision)
            return x * c["bn_scale"] + c["bn_shift"]
        return x

    def _param_device(self):
        for p in self.parameters():
            return p.device
        return None


class Attention_1(nn.Module):
    """The LAFF block (model/Attention.py:40-105): logits Linear(d -> 1), softmax over the L features, weighted sum
    (+ omega * mean-pool when with_ave), L2 normalise."""

    def __init__(self, embed_dim, with_ave=True, mul=False):
        super().__init__()
        self.with_ave = with_ave
        self.mul = mul
        self.embed_dim = embed_dim
        self.embedding_common = nn.Sequential(nn.Linear(embed_dim, 1))
        self.weights = 0
        self.global_emb_weight_net = nn.Linear(1, 1, False)
        self.change_raw_global_emb_weight(1)

    def get_raw_global_emb_weight(self):
        return self.global_emb_weight_net.weight.item()

    def change_raw_global_emb_weight(self, new_value: float):
        self.global_emb_weight_net.weight.data.fill_(new_value)

    def get_attention_weight(self):
        return torch.as_tensor(self.weights).clone().detach().cpu()

    def forward(self, local_embs: torch.Tensor, raw_global_emb=None):
        if raw_global_emb is not None:
            raise NotImplementedError("Attention_1 with an external raw_global_emb is not used by the LAFF configs")
        local_embs = local_embs.to(_cuda_device(self.embedding_common[0].weight.device)).float().contiguous()
        B, L, d = local_embs.shape
        srcs = [{"y": local_embs[:, l, :]} for l in range(L)]
        lin = self.embedding_common[0]
        out, _, att = ops.attention_pool(srcs, lin.weight.view(1, d), lin.bias.view(1), 1, d, self.with_ave, self.mul,
                                         omega=float(self.global_emb_weight_net.weight.item()) if self.with_ave else 0.0,
                                         want_att=True)
        self.weights = att.view(B, L)
        return out.view(B, d)


class Multi_head_MyApply_Attention(nn.Module):
    """H independent LAFF blocks on the H slices of the common space (model/Attention.py:473-552)."""

    def __init__(self, embed_dim, multi_heads=None, dim_per_head=None, with_ave=True, mul=True, split_head=True,
                 l2norm_each_head=False):
        super().__init__()
        if embed_dim is None:
            return
        if not split_head:
            raise NotImplementedError("split_head=False (every head sees the full vector) is not used by the LAFF configs")
        if l2norm_each_head:
            raise NotImplementedError("l2norm_each_head is off in every shipped config (base_config.py:125)")
        assert dim_per_head == embed_dim // multi_heads
        self.dim_per_head = dim_per_head
        self.multi_heads = multi_heads
        self.split_head = split_head
        self.with_ave, self.mul = with_ave, mul
        self.attention_layer = nn.Sequential()
        for i in range(multi_heads):
            self.attention_layer.add_module(str(i), Attention_1(dim_per_head, with_ave=with_ave, mul=mul))
        self.layer_norm = nn.LayerNorm(dim_per_head)  # present in the reference state_dict, unused in forward
        self.l2norm_each_head = l2norm_each_head
        self._cache = {}

    def head_params(self):
        ps = [self.attention_layer[h].embedding_common[0] for h in range(self.multi_heads)]
        key = (_PARAM_EPOCH,) + tuple((p.weight.data_ptr(), p.weight._version, p.bias._version) for p in ps)
        if self._cache.get("key") != key:
            w = torch.cat([p.weight.detach().view(1, -1) for p in ps], 0).float().contiguous()
            b = torch.cat([p.bias.detach().view(1) for p in ps], 0).float().contiguous()
            self._cache = {"key": key, "w": w, "b": b}
        return self._cache["w"], self._cache["b"]

    def pool(self, sources: Sequence[dict], out16_dtype=None, want_att=False):
        w, b = self.head_params()
        omega = float(self.attention_layer[0].global_emb_weight_net.weight.item()) if self.with_ave else 0.0
        out, out16, att = ops.attention_pool(sources, w, b, self.multi_heads, self.dim_per_head, self.with_ave, self.mul,
                                             omega=omega, out16_dtype=out16_dtype, want_att=want_att)
        if want_att:
            for h in range(self.multi_heads):
                self.attention_layer[h].weights = att[:, h, :]
        return out, out16

    def forward(self, local_embs, raw_global_emb=None, attn_mask=None):
        local_embs = local_embs.to(_cuda_device(self.layer_norm.weight.device)).float().contiguous()
        L = local_embs.shape[1]
        out, _ = self.pool([{"y": local_embs[:, l, :]} for l in range(L)], want_att=True)
        return out

    def get_raw_global_emb_weight(self):
        return self.attention_layer[0].global_emb_weight_net.weight.item()

    def change_raw_global_emb_weight(self, new_value: float):
        for i in range(self.multi_heads):
            self.attention_layer[i].global_emb_weight_net.weight.data.fill_(new_value)

    def get_attention_weight(self, head=0):
        return self.attention_layer[head].get_attention_weight().detach()


def get_attention_layer(attention_type: str, common_space_dim, encoder_num, opt, single_head_ok: bool = False):
    """model/model.py:95-208, restricted to the variants the LAFF / LAFF-ml scripts select.  The fusion
    nets pool over H heads and need the multi-head block; the single-head `Attention_1` names are only handed out
    where the caller says it can take them (the frame-level attention of LAFF-ml), so an unsupported configuration
    fails at construction rather than at the first encode."""
    if attention_type == "Multi_head_MyApply_Attention":
        return Multi_head_MyApply_Attention(
            common_space_dim, opt.multi_head_attention["heads"], common_space_dim // opt.multi_head_attention["heads"],
            with_ave=opt.attention_param_each_head["with_ave"], mul=opt.attention_param_each_head["mul"],
            split_head=opt.attention_param_each_head["split_head"], l2norm_each_head=getattr(opt, "attention_l2norm", False))
    table = {"attention_noAverageMul_Ave": (True, False), "attention_noAveNoAverageMul": (False, False),
             "attention_averageMul": (True, True), "average_AverageMul_noAve": (False, True)}
    if attention_type in table:
        if not single_head_ok:
            raise NotImplementedError("attention type %r (single-head Attention_1) as the fusion block: the shipped "
                                      "LAFF configs use 'Multi_head_MyApply_Attention' (configs/laff.py:41)" % attention_type)
        a, m = table[attention_type]
        return Attention_1(common_space_dim, with_ave=a, mul=m)
    raise NotImplementedError("attention type %r is an ablation variant outside the LAFF hot path" % attention_type)


_SINGLE_KERNEL = True  # use laff_fuse_forward when the configuration allows it (set False to force the two-kernel path)


def set_single_kernel_fusion(flag: bool) -> None:
    global _SINGLE_KERNEL
    _SINGLE_KERNEL = bool(flag)


# SMs that cast the next chunk's features while the fused kernel works on this one.  0 = serial, the default: measured
# (tools/bench_modeb.py, 1 M videos) serial 39.1 ms, overlapped on 8 / 12 / 16 / 24 SMs 43.4 / 41.4 / 41.8 / 41.4 ms -- the
# HBM-bound cast needs far more SMs than the fused kernel can spare (profiles/README.md).
_CAST_OVERLAP_SMS = 0
_CAST_STREAMS: Dict[int, tuple] = {}   # device index -> (cast stream, high-priority fused-kernel stream)


def set_cast_overlap(sms: int) -> None:
    global _CAST_OVERLAP_SMS
    _CAST_OVERLAP_SMS = max(0, int(sms))


def _fuse_single_kernel(features, attention, precision, out16_dtype, want_f32=True):
    """All projections + pooling in one kernel (csrc/fused.cu), `_FUSED_ROW_CHUNK` rows per launch so the 16-bit copies
    of the input features stay a bounded scratch (a 1 M-video gallery shard is fused in 8 launches).

    Experiment, off by default (set_cast_overlap): with more than one chunk the fp32 -> 16-bit casts of chunk c + 1 can
    run on a side stream, on `_CAST_OVERLAP_SMS` SMs, while the fused kernel of chunk c has the rest (double-buffered
    scratch, same results).  It measured slower than alternating the two, see the note at _CAST_OVERLAP_SMS."""
    B = features[0][0].shape[0]
    dev = features[0][0].device
    D = attention.multi_heads * attention.dim_per_head
    w, b = attention.head_params()
    prepared = [None if isinstance(x, ops.SparseRows) else tn.prepared(precision) for x, tn in features]
    out = torch.empty((B, D), dtype=torch.float32, device=dev) if want_f32 else None
    out16 = torch.empty((B, D), dtype=ops.torch_dtype(out16_dtype), device=dev) if out16_dtype is not None else None
    chunk = min(max(B, 1), _FUSED_ROW_CHUNK)
    n_chunks = (B + chunk - 1) // chunk if B > 0 else 0
    cast_ids = [i for i, (x, tn) in enumerate(features) if tn.fc1 is not None and not isinstance(x, ops.SparseRows)]
    overlap = n_chunks > 1 and precision != "bf16x3" and _CAST_OVERLAP_SMS > 0 and bool(cast_ids)
    scratch: Dict[tuple, torch.Tensor] = {}

    def cast_chunk(ci):
        """16-bit copies of chunk ci's projected features into scratch buffer set ci % 2 (one set when not overlapping)."""
        s, e = ci * chunk, min(B, (ci + 1) * chunk)
        outs = {}
        for i in cast_ids:
            x = features[i][0]
            key = (i, ci % 2 if overlap else 0)
            if n_chunks > 1 and key not in scratch:
                scratch[key] = torch.empty((chunk, (x.shape[1] + 7) // 8 * 8), dtype=_op_dtype(precision), device=dev)
            outs[i] = ops.cast_pad_16(x[s:e], _op_dtype(precision), out=scratch.get(key))
        return outs

    if overlap:
        # the fused kernel goes to a high-priority stream so that, when it and the next cast become runnable together,
        # its persistent CTAs get their SMs first and the cast's blocks land on the SMs it leaves free
        main = torch.cuda.current_stream(dev)
        di = dev.index if dev.index is not None else torch.cuda.current_device()
        if di not in _CAST_STREAMS:
            _CAST_STREAMS[di] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1))
        side, fstream = _CAST_STREAMS[di]
        total_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        side_sms = min(_CAST_OVERLAP_SMS, total_sms // 4)
        side_sms -= side_sms % 2
        cast_done, fuse_done = {}, {}
        casted = {0: cast_chunk(0)}                       # the first chunk has nothing to hide behind: whole device
        cast_done[0] = torch.cuda.Event()
        cast_done[0].record(main)
        side.wait_stream(main)
        fstream.wait_stream(main)
    for ci in range(n_chunks):
        s, e = ci * chunk, min(B, (ci + 1) * chunk)
        x16s = None
        if overlap:
            fstream.wait_event(cast_done[ci])
            x16s = casted.pop(ci)
        fc, tiled = [], []
        for i, ((x, tn), c) in enumerate(zip(features, prepared)):
            xs = x[s:e]
            if isinstance(xs, ops.SparseRows):
                # sparse BoW: gather-sum projection (bias, activation, BN applied), then a "tiled" feature of full width
                tiled.append({"x": tn.project_sparse(xs), "bn_scale": None, "bn_shift": None})
            elif tn.fc1 is not None:
                if precision == "bf16x3":
                    x16 = ops.split3_16(xs, 0, torch.bfloat16)
                elif x16s is not None:
                    x16 = x16s[i]
                else:
                    key = (i, 0)
                    if n_chunks > 1 and key not in scratch:
                        scratch[key] = torch.empty((chunk, (x.shape[1] + 7) // 8 * 8), dtype=_op_dtype(precision), device=dev)
                    x16 = ops.cast_pad_16(xs, _op_dtype(precision), out=scratch.get(key))
                fc.append({"x16": x16, "w16": c["w16"], "bias": c["bias"], "activation": tn.activation_name,
                           "bn_scale": c.get("bn_scale"), "bn_shift": c.get("bn_shift")})
            else:
                tiled.append({"x": xs, "bn_scale": c.get("bn_scale"), "bn_shift": c.get("bn_shift")})
        # an enclosing SM budget (the retrieval pipeline's side stages) stays in force unless this loop sets its own
        with (torch.cuda.stream(fstream) if overlap else contextlib.nullcontext()), \
                (ops.sm_limit(total_sms - side_sms) if overlap and ci + 1 < n_chunks else contextlib.nullcontext()):
            ops.fuse_forward(fc, tiled, w, b, attention.multi_heads, attention.dim_per_head, want_f32=False,
                             out=None if out is None else out[s:e], out16=None if out16 is None else out16[s:e])
        if overlap:
            fuse_done[ci] = torch.cuda.Event()
            fuse_done[ci].record(fstream)
            if ci + 1 < n_chunks:                         # issued AFTER the fused kernel of this chunk
                with torch.cuda.stream(side), ops.sm_limit(side_sms):
                    if ci - 1 in fuse_done:               # the buffer set of chunk ci + 1 was read by the fused kernel of ci - 1
                        side.wait_event(fuse_done[ci - 1])
                    casted[ci + 1] = cast_chunk(ci + 1)
                    cast_done[ci + 1] = torch.cuda.Event()
                    cast_done[ci + 1].record(side)
    if overlap:
        main.wait_stream(fstream)
        main.wait_stream(side)
        for t in list(scratch.values()) + [t for t in (out, out16) if t is not None]:   # used on streams they were not allocated on
            t.record_stream(side)
            t.record_stream(fstream)
            t.record_stream(main)
    H, dh = attention.multi_heads, attention.dim_per_head
    return (None if out is None else out.view(B, H, dh)), (None if out16 is None else out16.view(B, H, dh))


def _single_kernel_ok(features, attention, want_att) -> bool:
    if not _SINGLE_KERNEL or want_att or attention.with_ave or attention.mul or attention.dim_per_head != 512:
        return False
    n_fc = sum(1 for x, tn in features if tn.fc1 is not None and not isinstance(x, ops.SparseRows))
    n_tiled = len(features) - n_fc
    if not (1 <= n_fc <= 4 and n_tiled <= 2):
        return False
    D = attention.multi_heads * 512
    return all(tn.fc1 is not None or (x.shape[1] % 128 == 0 and D % x.shape[1] == 0) for x, tn in features)


def _fuse(features: Sequence, attention: Multi_head_MyApply_Attention, device, precision, out16_dtype=None,
          want_att=False, want_f32=True):
    """features: list of (x fp32 [B, d] device tensor, TransformNet).  One fused kernel when the configuration allows
    (head_dim 512, with_ave = mul = False, attention weights not requested); otherwise projection GEMMs + pooling kernel,
    chunked by rows."""
    if _single_kernel_ok(features, attention, want_att):
        return _fuse_single_kernel(features, attention, precision, out16_dtype, want_f32 or out16_dtype is None)
    B = features[0][0].shape[0]
    D = attention.multi_heads * attention.dim_per_head
    outs, outs16 = [], []
    scratch: Dict[int, torch.Tensor] = {}
    for s in range(0, max(B, 1), _ROW_CHUNK):
        e = min(B, s + _ROW_CHUNK)
        srcs = []
        for i, (x, tn) in enumerate(features):
            xs = x[s:e]
            if tn.fc1 is not None:
                buf = scratch.get(i)
                if buf is None or buf.shape[0] < e - s:
                    buf = torch.empty((min(_ROW_CHUNK, B), D), dtype=torch.float32, device=device)
                    scratch[i] = buf
                srcs.append({"y": tn.project(xs, precision, out=buf[: e - s])})
            else:
                c = tn.prepared(precision)
                srcs.append({"x": xs, "bn_scale": c.get("bn_scale"), "bn_shift": c.get("bn_shift")})
        o, o16 = attention.pool(srcs, out16_dtype=out16_dtype, want_att=B <= _ROW_CHUNK)  # Attention_1.weights, as the reference keeps
        outs.append(o)
        outs16.append(o16)
    out = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
    out16 = None if out16_dtype is None else (outs16[0] if len(outs16) == 1 else torch.cat(outs16, 0))
    return out, out16


class VisMutiTransformNet(nn.Module):
    """dict of video features -> dict of features in the common space (model/model.py:1787-1827)."""

    def __init__(self, opt, space_dict: dict):
        super().__init__()
        if opt is None:
            return
        self.opt = opt
        self.vis_net_space_dict = space_dict
        self.common_space_dim = opt.vis_fc_layers[1]
        for each in space_dict.keys():
            if each not in opt.vis_no_transform:
                self.add_module(each, TransformNet((space_dict[each], opt.vis_fc_layers[1]), opt))
            else:
                self.add_module(each, TransformNet((space_dict[each], opt.vis_fc_layers[1]), None, dropout=None,
                                                   batch_norm=True, activation=False, fc=False))

    def forward(self, vis_input, txt_emb=None, vis_frame_feat_dict_input=None):
        out = {}
        mods = dict(self.named_children())
        heads = self.opt.multi_head_attention["heads"]
        for name in self.vis_net_space_dict.keys():
            x = vis_input[name]
            if name in self.opt.vis_no_transform:
                x = x.repeat(1, heads)
            out[name] = mods[name](x)
        return out


class VisMutiTransformNetAddAttnetion(nn.Module):
    """dict of video features -> [B, H, d_h] multi-space embedding (model/model.py:1830-1881)."""

    def __init__(self, opt, space_dict: dict):
        super().__init__()
        if opt is None:
            return
        if getattr(opt, "vis_expert_embedding", {"expert": False}).get("expert") or \
                getattr(opt, "vis_expert_embedding", {"l2norm": False}).get("l2norm"):
            raise NotImplementedError("expert embeddings are off in the shipped configs (base_config.py:158)")
        self.opt = opt
        self.vis_net_space_dict = space_dict
        self.common_space_dim = opt.vis_fc_layers[1]
        self.VisMutiTransformNet = VisMutiTransformNet(opt, space_dict)
        self.attention_layer = get_attention_layer(opt.vis_attention, self.common_space_dim, len(space_dict), opt)
        self.expert_embedding = None

    def encode(self, vis_input, out16_dtype=None, precision=None, want_att=False, want_f32=True):
        """-> (fp32 [B, H, d_h] embeddings, 16-bit copy or None).  want_f32 = False with an out16_dtype skips the fp32
        copy where the single-kernel path runs (gallery encoding keeps only the 16-bit rows)."""
        precision = precision or _loss.get_precision()
        dev = _cuda_device(self.attention_layer.layer_norm.weight.device)
        if self.training:
            raise NotImplementedError("train-mode forward of the fusion net is SURVEY §8f N4; call .eval()")
        mods = dict(self.VisMutiTransformNet.named_children())
        feats = [(vis_input[name].to(dev, non_blocking=True).float(), mods[name]) for name in self.vis_net_space_dict.keys()]
        return _fuse(feats, self.attention_layer, dev, precision, out16_dtype, want_att, want_f32)

    def forward(self, vis_input, txt_emb=None, vis_frame_feat_dict_input=None):
        return self.encode(vis_input)[0]

    def get_attention_weight(self, vis_input, txt_emb=None):
        self.encode(vis_input, want_att=True)
        return self.attention_layer.get_attention_weight()


# ----------------------------------------------------------------------------------------------------------------
# text encoders: strings -> per-encoder features on the device
# ----------------------------------------------------------------------------------------------------------------
class TxtEncoder(nn.Module):
    """model/model.py:311-320."""

    def __init__(self, opt):
        super().__init__()

    def forward(self, caption_feat_dict, task3=False):
        return {"text_features": caption_feat_dict["caption"]}


class GruTxtEncoder(TxtEncoder):
    """model/model.py:322-387: IndexVec token ids -> nn.Embedding -> 1-layer GRU -> mean / last / mean_last pooling.
    Parameters live in `we` (nn.Embedding) and `rnn` (nn.GRU) so reference checkpoints load by name; the arithmetic is
    laff_b200.text.gru_encode (device kernels), not torch's GRU."""

    def __init__(self, opt):
        super().__init__(opt)
        self.bigru = False
        if int(opt.rnn_layer) != 1:
            raise NotImplementedError("rnn_layer = %s: the shipped configs use one GRU layer (base_config.py:39)" % opt.rnn_layer)
        self.pooling = opt.pooling
        self.rnn_size = int(opt.rnn_size)
        self.t2v_idx = opt.t2v_idx
        self.we = nn.Embedding(len(self.t2v_idx.vocab), int(opt.we_dim))
        if int(opt.we_dim) == 500 and getattr(opt, "we", None) is not None:
            self.we.weight = nn.Parameter(torch.as_tensor(opt.we, dtype=torch.float32))  # pre-trained 500-d w2v
        self.rnn = nn.GRU(int(opt.we_dim), self.rnn_size, 1, batch_first=True, bidirectional=False)
        self._prepared: Dict[str, torch.Tensor] = {}

    def train(self, mode: bool = True):
        self._prepared = {}
        return super().train(mode)

    def load_state_dict(self, *a, **k):
        self._prepared = {}
        return super().load_state_dict(*a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._prepared = {}
        return super()._load_from_state_dict(*a, **k)

    def forward(self, caption_feat_dict, task3=False):
        from . import text as _text
        if self.training:
            raise NotImplementedError("in train mode the GRU runs inside the model's training step (model(train_data)), which "
                                      "keeps what backward-through-time needs; call .eval() for a plain forward")
        ids, lengths = self.t2v_idx.encoding_batch(caption_feat_dict["caption"])
        dev = _cuda_device(self.we.weight.device)
        out = _text.gru_encode(self.we.weight, self.rnn.weight_ih_l0, self.rnn.weight_hh_l0, self.rnn.bias_ih_l0,
                               self.rnn.bias_hh_l0, torch.from_numpy(ids).to(dev), torch.from_numpy(lengths).to(dev),
                               self.pooling, self._prepared)
        return {"text_features": out}


class BoWTxtEncoder(TxtEncoder):
    """model/model.py:399-416."""

    def __init__(self, opt):
        super().__init__(opt)
        self.t2v_bow = opt.t2v_bow

    def forward(self, caption_feat_dict, task3=False, sparse=False):
        """Dense count vectors like the reference; sparse=True (the eval-mode fusion path asks for it) keeps the CSR
        token ids, which the projection consumes directly (ops.bow_project) -- no [B, |vocab|] matrix is built."""
        if sparse:
            return {"text_features": self.t2v_bow.encode_sparse(caption_feat_dict["caption"])}
        return {"text_features": self.t2v_bow.encode_batch(caption_feat_dict["caption"])}


class W2VTxtEncoder(TxtEncoder):
    """model/model.py:419-434."""

    def __init__(self, opt):
        super().__init__(opt)
        self.t2v_w2v = opt.t2v_w2v

    def forward(self, caption_feat_dict, task3=False):
        return {"text_features": self.t2v_w2v.encode_batch(caption_feat_dict["caption"])}


# reference encoder name -> (feature key accepted in caption_feat_dict, alternatives)
_TXT_ENCODERS = (("rnn_encoder", ("gru", "rnn_encoder")), ("bow_encoder", ("bow", "bow_encoder")),
                 ("w2v_encoder", ("w2v", "w2v_encoder")), ("CLIP_encoder", ("clip", "CLIP_encoding", "CLIP_encoder")))


class MultiScaleTxtEncoderAttention(nn.Module):
    """Per-encoder text features -> [B, H, d_h] (model/model.py:1641-1709, init_transform :622-681).

    At this tier the encoder outputs are inputs: ``caption_feat_dict`` carries 'gru' [B, rnn_size], 'bow' [B, |V|],
    'w2v' [B, 500] and 'clip' / 'CLIP_encoding' [B, 512] tensors (the reference reads CLIP text features precomputed
    when frozen, model/model.py:497-498).  Encoder order = the reference's encoder_name_list (rnn, bow, w2v, CLIP).
    """

    def __init__(self, opt):
        super().__init__()
        if getattr(opt, "txt_expert_embedding", {"expert": False}).get("expert") or \
                getattr(opt, "txt_expert_embedding", {"l2norm": False}).get("l2norm"):
            raise NotImplementedError("expert embeddings are off in the shipped configs")
        self.opt = opt
        te = opt.text_encoding
        D = opt.txt_fc_layers[1]
        self.space_dict = {}
        if te["rnn_encoding"]["name"].split("_", 1)[0] == "gru":
            self.space_dict["rnn_encoder"] = opt.rnn_size
        elif te["rnn_encoding"]["name"].split("_", 1)[0] == "bigru":
            self.space_dict["rnn_encoder"] = opt.rnn_size * 2
        if "no" not in te["bow_encoding"]["name"]:
            self.space_dict["bow_encoder"] = opt.t2v_bow.ndims
        if "no" not in te["w2v_encoding"]["name"]:
            self.space_dict["w2v_encoder"] = opt.t2v_w2v.ndims
        if "no" not in te["CLIP_encoding"]["name"]:
            self.space_dict["CLIP_encoder"] = opt.clip_opt["size"]
        self.encoder_name_list = [n for n, _ in _TXT_ENCODERS if n in self.space_dict]
        self.txt_encoder_num = len(self.encoder_name_list)
        # String front-end (init_txt_encoder, model/model.py:558-613): built when the config carries the vocabulary
        # objects; otherwise the per-encoder features must arrive precomputed in caption_feat_dict.
        self.encoder = nn.Module()
        if "rnn_encoder" in self.space_dict and hasattr(getattr(opt, "t2v_idx", None), "vocab"):
            if te["rnn_encoding"]["name"].split("_", 1)[0] != "gru":
                raise NotImplementedError("bigru text encoder: the shipped LAFF configs use gru_mean (base_config.py:21)")
            opt.pooling = te["rnn_encoding"]["name"].split("_", 1)[1]
            self.encoder.add_module("rnn_encoder", GruTxtEncoder(opt))
        if "bow_encoder" in self.space_dict and hasattr(opt.t2v_bow, "encode_batch"):
            self.encoder.add_module("bow_encoder", BoWTxtEncoder(opt))
        if "w2v_encoder" in self.space_dict and hasattr(opt.t2v_w2v, "encode_batch"):
            self.encoder.add_module("w2v_encoder", W2VTxtEncoder(opt))
        self.transform_layer = nn.Module()
        # registration order follows init_transform (rnn, w2v, bow, CLIP) so that state_dict() order matches
        for name in ("rnn_encoder", "w2v_encoder", "bow_encoder", "CLIP_encoder"):
            if name not in self.space_dict:
                continue
            if name == "CLIP_encoder":
                co = opt.clip_opt
                if "CLIP_encoder" in opt.txt_no_transform:
                    tn = TransformNet((co["size"], D), None, co["transform_dropout"], co["transform_batch_norm"], False, False)
                else:
                    tn = TransformNet((co["size"], D), None, co["transform_dropout"], co["transform_batch_norm"],
                                      co["transform_activation"])
            else:
                tn = TransformNet((self.space_dict[name], D), None, opt.dropout, opt.batch_norm, opt.activation)
            self.transform_layer.add_module(name + "_transform", tn)
        self.attention_layer = get_attention_layer(opt.txt_attention, D, self.txt_encoder_num, opt)

    def _feature(self, caption_feat_dict, enc, sparse_ok=False):
        """The feature of one encoder: precomputed under one of its keys, or from the caption strings.  The BoW feature
        may arrive sparse ('bow_csr': ops.SparseRows or an (offsets, ids) pair); sparse_ok says the caller can consume
        it that way, otherwise it is expanded to the dense count matrix the reference builds (txt2vec.py:56-63)."""
        if enc == "bow_encoder" and "bow_csr" in caption_feat_dict:
            x = caption_feat_dict["bow_csr"]
            if not isinstance(x, ops.SparseRows):
                x = ops.SparseRows(torch.as_tensor(x[0]), torch.as_tensor(x[1]), self.space_dict["bow_encoder"])
            if x.ndims != self.space_dict["bow_encoder"]:
                raise ops.LaffError("bow_csr: vocabulary size %d, the model expects %d" % (x.ndims, self.space_dict["bow_encoder"]))
            return x if sparse_ok else x.to(_cuda_device(self.attention_layer.layer_norm.weight.device)).dense()
        for key in dict(_TXT_ENCODERS)[enc]:
            if key in caption_feat_dict:
                return caption_feat_dict[key]
        front = dict(self.encoder.named_children()).get(enc)
        if front is not None and "caption" in caption_feat_dict:
            if enc == "bow_encoder" and sparse_ok:
                return front(caption_feat_dict, sparse=True)["text_features"]
            return front(caption_feat_dict)["text_features"]
        raise KeyError("caption_feat_dict has no feature for %s (expected one of %s)" % (enc, dict(_TXT_ENCODERS)[enc]))

    def encode(self, caption_feat_dict, out16_dtype=None, precision=None, want_att=False, want_f32=True):
        precision = precision or _loss.get_precision()
        dev = _cuda_device(self.attention_layer.layer_norm.weight.device)
        if self.training:
            raise NotImplementedError("train-mode forward of the fusion net is SURVEY §8f N4; call .eval()")
        mods = dict(self.transform_layer.named_children())
        feats = [(self._feature(caption_feat_dict, n, sparse_ok=True).to(dev, non_blocking=True).float(), mods[n + "_transform"])
                 for n in self.encoder_name_list]
        return _fuse(feats, self.attention_layer, dev, precision, out16_dtype, want_att, want_f32)

    def forward(self, caption_feat_dict, visual_emb=None, task3=False):
        return self.encode(caption_feat_dict)[0]

    def get_attention_weight(self, caption_feat_dict, visual_emb=None):
        self.encode(caption_feat_dict, want_att=True)
        return self.attention_layer.get_attention_weight()


class VisMutiTransformNetPlusFrameFeat(nn.Module):
    """LAFF-ml video side: frame-level LAFF per video, then LAFF over [video-level features..., pooled frame feature]
    (model/model.py:2101-2194)."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        space_dict = opt.vis_fc_layers[0]
        self.vis_net_space_dict = space_dict
        for each in space_dict.keys():
            if each not in opt.vis_no_transform:
                self.add_module(each, TransformNet((space_dict[each], opt.vis_fc_layers[1]), opt))
            else:
                self.add_module(each, TransformNet((space_dict[each], opt.vis_fc_layers[1]), None, dropout=None,
                                                   batch_norm=True, activation=False, fc=False))
        self.vis_attention_layer = get_attention_layer(opt.vis_attention, opt.vis_fc_layers[1], len(space_dict), opt)
        self.frame_attention = nn.ModuleDict()
        for each in opt.vid_frame_feats:
            if opt.vis_frame_addFC:
                raise NotImplementedError("vis_frame_addFC=True is not used by the LAFF-ml script (FrameLaff...:58)")
            self.frame_attention[each] = nn.Sequential(
                get_attention_layer(opt.vis_frame_attention, opt.vis_fc_layers[0][each], 1, opt, single_head_ok=True))

    def frame_pool(self, feat_name, frames: torch.Tensor) -> torch.Tensor:
        att: Attention_1 = self.frame_attention[feat_name][0]
        lin = att.embedding_common[0]
        return ops.frame_pool(frames, lin.weight, float(lin.bias.item()), att.with_ave, att.mul,
                              omega=float(att.global_emb_weight_net.weight.item()) if att.with_ave else 0.0)

    def encode(self, vis_input, vis_frame_feat_dict_input, out16_dtype=None, precision=None, want_f32=True):
        precision = precision or _loss.get_precision()
        dev = _cuda_device(self.vis_attention_layer.layer_norm.weight.device)
        if self.training:
            raise NotImplementedError("train-mode forward of the fusion net is SURVEY §8f N4; call .eval()")
        feats_in = dict(vis_input) if self.opt.frame_feat_with_video_feat else {}
        for feat_name, fr in vis_frame_feat_dict_input.items():
            if feat_name == "mask_tensor":
                continue  # the reference's mask slice is a no-op: zero-padded frames take part
            feats_in[feat_name] = self.frame_pool(feat_name, fr.to(dev, non_blocking=True).float())
        mods = dict(self.named_children())
        feats = [(x.to(dev, non_blocking=True).float(), mods[name]) for name, x in feats_in.items()]
        return _fuse(feats, self.vis_attention_layer, dev, precision, out16_dtype, False, want_f32)

    def forward(self, vis_input, vis_frame_feat_dict_input, txt_emb=None):
        return self.encode(vis_input, vis_frame_feat_dict_input)[0]


class W2VVPP(nn.Module):
    """Shared model surface (model/model.py:751-1128): similarity, loss, predict."""

    def _init_vis_net(self, opt):
        raise NotImplementedError

    def _init_txt_net(self, opt):
        self.txt_net = MultiScaleTxtEncoderAttention(opt)

    def __init__(self, opt):
        super().__init__()
        if opt is None:
            return
        self._init_vis_net(opt)
        self._init_txt_net(opt)
        self.opt = opt
        self.grad_clip = getattr(opt, "grad_clip", 2)
        kind = getattr(opt, "loss", "mrl")
        if kind == "mrl":
            self.criterion = MarginRankingLoss(margin=opt.margin, measure=opt.measure, max_violation=opt.max_violation,
                                               cost_style=opt.cost_style, direction=opt.direction)
        elif kind == "dsl":
            self.criterion = DualSoftmaxLoss()                                          # model/model.py:1995-1996
        else:
            raise NotImplementedError("loss %r: the reference's CELoss is dead code (its cal_loss signature does not match "
                                      "its call, loss.py:275-285); 'mrl' and 'dsl' are built" % kind)
        self.criterion_with_score = MarginRankingLossWithScore(margin=opt.margin, max_violation=opt.max_violation,
                                                               cost_style=opt.cost_style, direction=opt.direction)
        self.iters = 0

    # ---------------------------------------------------------------- similarity (model/model.py:1003-1016, 1567-1578)
    @staticmethod
    def compute_sim(query_embs, retro_embs, measure="cosine", device=None):
        if measure == "cosine":
            return _loss.cosine_sim(query_embs, retro_embs)
        elif measure == "hist":
            raise Exception("measure 'hist' is outside the LAFF hot path")
        elif measure == "euclidean":
            raise Exception("Not implemented")
        else:
            raise Exception("%s is invalid" % measure)

    def get_txt2vis_matrix(self, txt_embs, vis_embs, measure="cosine"):
        if txt_embs.dim() == vis_embs.dim() == 2:
            return self.compute_sim(txt_embs, vis_embs, measure)
        if measure != "cosine":
            self.compute_sim(txt_embs[:, 0, :], vis_embs[:, 0, :], measure)  # raises like the reference
        H = vis_embs.shape[1]
        q16, g16 = _loss._operands(txt_embs.reshape(txt_embs.shape[0], -1), vis_embs.reshape(vis_embs.shape[0], -1), H,
                                   _loss.get_precision())
        return ops.sim_dense(q16, g16, 1.0 / H)  # mean over heads of the per-head cosine

    # ---------------------------------------------------------------- loss (model/model.py:840-865, 2032-2048)
    def compute_loss(self, vis_embs, txt_embs, vis_embs_multi_labels=0, txt_embs_multi_labels=0, labels_embs=0):
        if vis_embs.dim() == txt_embs.dim() == 2 or vis_embs.dim() == txt_embs.dim() == 3:
            loss = self.criterion(txt_embs, vis_embs)  # 3-D: the kernel sums over heads
            return loss, {"triplet_loss": loss}
        raise Exception("vis_embs dims are not equal to txt_embs dims")

    # ---------------------------------------------------------------- training step (model/model.py:964-1001)
    def _stage_train_inputs(self, train_data, dev):
        """Device fp32 tensors of one batch: ({text encoder name: feature}, {video feature name: feature}).  String
        front-ends (BoW / word2vec) run here, before anything that a CUDA graph captures."""
        caps = train_data["captions"]
        fronts = dict(self.txt_net.encoder.named_children())
        txt = {}
        for n in self.txt_net.encoder_name_list:
            precomputed = any(k in caps for k in dict(_TXT_ENCODERS)[n])
            if n == "rnn_encoder" and not precomputed and "rnn_encoder" in fronts and "caption" in caps:
                # the GRU front-end trains with the model: token ids go to the device, the recurrence and its backward
                # through time run inside the step (laff_b200.text.gru_encode_train / gru_backward)
                ids, lengths = fronts["rnn_encoder"].t2v_idx.encoding_batch(caps["caption"])
                txt["rnn_ids"] = torch.from_numpy(ids).to(dev, non_blocking=True)
                txt["rnn_len"] = torch.from_numpy(lengths).to(dev, non_blocking=True)
                continue
            txt[n] = self.txt_net._feature(caps, n).to(dev, non_blocking=True).float()
        if isinstance(self.vis_net, VisMutiTransformNetPlusFrameFeat):
            frames = train_data.get("vis_frame_feat_dict") or {}
            names = [n for n in self.vis_net.vis_net_space_dict.keys() if n not in self.vis_net.frame_attention]
            if not self.opt.frame_feat_with_video_feat:
                names = []
            vis = {n: train_data["vis_feats"][n].to(dev, non_blocking=True).float() for n in names}
            for n in self.vis_net.frame_attention.keys():  # zero-padded frames take part, as in the reference
                vis["frames/" + n] = frames[n].to(dev, non_blocking=True).float().contiguous()
            return txt, vis
        vis = {n: train_data["vis_feats"][n].to(dev, non_blocking=True).float() for n in self.vis_net.vis_net_space_dict.keys()}
        return txt, vis

    def _train_step_device(self, txt, vis, precision):
        """forward (train mode) -> loss -> backward -> clip + optimizer step, no host sync.  The text net and the
        video net are independent until the loss and again after it, so the text side runs on a second stream (forked
        from / joined into the current one with events: inside a CUDA-graph capture these become two parallel branches
        of the graph) -- at B = 128 every kernel fills only a fraction of the SMs."""
        from .train import FusionTrainStep
        self._seed_dev.add_(1)
        dev = self._seed_dev.device
        main = torch.cuda.current_stream(dev)
        side = getattr(self, "_side_stream", None)
        if side is None or side.device != dev:
            side = self._side_stream = torch.cuda.Stream(dev)
        outs = {}

        def run_step(key, feats, att):
            step = self._steps.get(key)
            if step is None or step.att is not att or step.precision != precision:
                step = self._steps[key] = FusionTrainStep(att, precision)
            outs[key] = step.forward(feats, self._seed_base * 2 + (key == "vis"), self._seed_dev)

        # ---- forward: text side on the second stream ----
        gru_cache, gru_idx = None, None
        side.wait_stream(main)
        with torch.cuda.stream(side):
            tmods = dict(self.txt_net.transform_layer.named_children())
            tfeats = []
            for n in self.txt_net.encoder_name_list:
                if n == "rnn_encoder" and "rnn_ids" in txt:
                    from . import text as _text
                    enc = dict(self.txt_net.encoder.named_children())["rnn_encoder"]
                    feat, gru_cache = _text.gru_encode_train(enc.we.weight, enc.rnn.weight_ih_l0, enc.rnn.weight_hh_l0, enc.rnn.bias_ih_l0,
                                                             enc.rnn.bias_hh_l0, txt["rnn_ids"], txt["rnn_len"], enc.pooling)
                    gru_idx = len(tfeats)
                    tfeats.append((feat, tmods[n + "_transform"], True))
                else:
                    tfeats.append((txt[n], tmods[n + "_transform"]))
            run_step("txt", tfeats, self.txt_net.attention_layer)
        frame_ml = isinstance(self.vis_net, VisMutiTransformNetPlusFrameFeat)
        pooled = {}
        if frame_ml:
            # LAFF-ml (model/model.py:2147-2190): frame-level Attention_1 per video, its output joins the video-level
            # features as a no-transform feature.  (The logit bias cancels in the softmax: passed as 0, no host read.)
            vmods = dict(self.vis_net.named_children())
            vis_att = self.vis_net.vis_attention_layer
            vfeats = []
            for n, x in vis.items():
                if n.startswith("frames/"):
                    fa = self.vis_net.frame_attention[n[7:]][0]
                    if fa.with_ave or fa.mul:
                        raise NotImplementedError("training the mean-residual / product frame attention is not built")
                    pooled[len(vfeats)] = (n[7:], x)
                    vfeats.append((ops.frame_pool(x, fa.embedding_common[0].weight, 0.0), vmods[n[7:]], True))
                else:
                    vfeats.append((x, vmods[n]))
        else:
            vmods = dict(self.vis_net.VisMutiTransformNet.named_children())
            vis_att = self.vis_net.attention_layer
            # train-mode quirk of the reference: an all-zero video feature becomes noise (model/model.py:1819-1821).
            # Selected on the device: the reference's `torch.nonzero(...)` costs a host synchronisation per feature per step.
            vfeats = [(torch.where((x != 0).any(), x, torch.randn_like(x)), vmods[n]) for n, x in vis.items()]
        run_step("vis", vfeats, vis_att)
        main.wait_stream(side)
        # ---- loss (both embeddings) ----
        c = self.criterion
        if isinstance(c, DualSoftmaxLoss):
            loss, d_txt, d_vis = ops.dsl_forward_backward(outs["txt"], outs["vis"], 1000.0)
        else:
            loss, d_txt, d_vis = ops.mrl_forward_backward(outs["txt"], outs["vis"], c.margin, c.max_violation, c.direction, c.cost_style)
        # ---- backward: text side (and the GRU's backward through time) on the second stream ----
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dxt = self._steps["txt"].backward(d_txt)
            if gru_cache is not None:
                from . import text as _text
                from .train import _grad_buffer
                enc = dict(self.txt_net.encoder.named_children())["rnn_encoder"]
                grads = {"we": _grad_buffer(enc.we.weight), "w_ih": _grad_buffer(enc.rnn.weight_ih_l0), "w_hh": _grad_buffer(enc.rnn.weight_hh_l0),
                         "b_ih": _grad_buffer(enc.rnn.bias_ih_l0), "b_hh": _grad_buffer(enc.rnn.bias_hh_l0)}
                _text.gru_backward(gru_cache, dxt[gru_idx], enc.we.weight, enc.rnn.weight_ih_l0, enc.rnn.weight_hh_l0, grads)
        dxs = self._steps["vis"].backward(d_vis)
        for idx, (name, frames) in pooled.items():
            lin = self.vis_net.frame_attention[name][0].embedding_common[0]
            bufs = self._frame_grads.get(name)
            if bufs is None:
                bufs = self._frame_grads[name] = (torch.zeros_like(lin.weight), torch.zeros_like(lin.bias))
            lin.weight.grad, lin.bias.grad = bufs
            ops.frame_pool_backward(frames, lin.weight, dxs[idx], bufs[0], bufs[1])
        main.wait_stream(side)
        self.last_grad_norm = self.optimizer.step()
        return loss

    def _make_optimizer(self):
        from .train import DeviceGradScaler, DeviceOptimizer
        opt = self.opt
        kind = getattr(opt, "optimizer", "rmsprop")
        eps = self._adam_eps if kind == "adam" else None
        # config.float16: the reference trains under autocast + GradScaler and clips the *scaled* gradients
        # (model/model.py:970-989); `self.scaler` mirrors its attribute of the same name (model/model.py:793)
        self.scaler = DeviceGradScaler() if getattr(opt, "float16", False) else None
        return DeviceOptimizer(list(self.parameters()), kind=kind, lr=getattr(opt, "lr", 1e-4), eps=eps,
                               max_grad_norm=self.grad_clip if self.grad_clip and self.grad_clip > 0 else 0.0, scaler=self.scaler)

    _adam_eps = 1e-8  # torch default (model/model.py:824); the LAFF class overrides it with 1e-4 (model/model.py:2022)
    use_cuda_graph = True   # replay the whole step as one CUDA graph from the 4th step on (launch-bound at B = 128)
    train_precision = "bf16x3"

    def forward(self, train_data, epoch=None):
        """One training step (model/model.py:964-1001): forward of both nets in train mode, summed per-head
        MarginRankingLoss, backward, clip_grad_norm_(params, grad_clip), optimizer step.  Returns loss_items.
        Everything runs through the C ABI (laff_b200/train.py); no autograd graph is built.  After three eager steps
        the step is captured once into a CUDA graph and replayed (same shapes): ~100 launches become one."""
        global _PARAM_EPOCH
        opt = self.opt
        if getattr(opt, "negative", False):
            raise NotImplementedError("negation-aware training (cal_foward_neg) is outside the LAFF hot path")
        if not getattr(opt, "multi_space", True):
            raise NotImplementedError("training through the score-matrix loss branch (multi_space = False) is not built")
        if not self.training:
            raise ops.LaffError("model(train_data) is a training step: call model.train() first")
        self.iters += 1
        dev = _cuda_device(next(self.parameters()).device)
        # config.float16 (AMP in the reference, model/model.py:970-989): fp16 tensor-core operands with fp32 accumulation,
        # master weights and gradients.  Nothing is stored in fp16 here, but the reference's ordering -- clip_grad_norm_ on
        # the scaled gradients, skipped steps, the moving loss scale -- is observable in the trained weights, so the
        # optimizer reproduces it on the device (train.DeviceGradScaler / laff_optimizer_step_scaled).
        precision = "fp16" if getattr(opt, "float16", False) else self.train_precision
        if getattr(self, "optimizer", None) is None:
            self.optimizer = self._make_optimizer()
            self._steps = {}
            self._seed_base = int(getattr(opt, "seed", 0) or 0) << 20
            self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self._frame_grads = {}
            self._graph, self._graphs = None, {}
        with ops.nvtx_range("train_stage_inputs"):
            txt, vis = self._stage_train_inputs(train_data, dev)
        sig = (precision,) + tuple((k, tuple(v.shape)) for k, v in list(txt.items()) + list(vis.items()))
        g = self._graphs.get(sig) if self._graphs else None
        self._graph = g
        if g is not None:
            self.optimizer.sync_lr()
            for k, v in txt.items():
                g["txt"][k].copy_(v, non_blocking=True)
            for k, v in vis.items():
                g["vis"][k].copy_(v, non_blocking=True)
            with ops.nvtx_range("train_step_graph_replay"):
                g["graph"].replay()
            loss = g["loss"].clone()
        else:
            with ops.nvtx_range("train_step_eager"):
                loss = self._train_step_device(txt, vis, precision)
            if self.use_cuda_graph and self.iters >= 3 and all(st.capturable for st in self._steps.values()):
                if len(self._graphs) >= 16:  # caption lengths vary: keep the graphs of the most recent shapes
                    self._graphs.pop(next(iter(self._graphs)))
                self._graph = self._graphs[sig] = self._capture_train_graph(txt, vis, precision, sig)
        _PARAM_EPOCH += 1  # parameters changed behind torch's version counters: drop the eval-mode operand caches
        return {"triplet_loss": loss}

    def _capture_train_graph(self, txt, vis, precision, sig):
        st_txt = {k: v.clone() for k, v in txt.items()}
        st_vis = {k: v.clone() for k, v in vis.items()}
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph):
            loss = self._train_step_device(st_txt, st_vis, precision)
        return {"graph": graph, "txt": st_txt, "vis": st_vis, "loss": loss, "sig": sig}

    def train(self, mode: bool = True):
        self._graph, self._graphs = None, {}  # BatchNorm / dropout behaviour is baked into a captured step
        return super().train(mode)

    def _encode_vis(self, output_dict, out16_dtype=None):
        return self.vis_net.encode(output_dict["vis_feat_dict"], out16_dtype)

    def predict(self, txt_loader, vis_loader, measure, record_emb=False):
        """Dense score matrix for small galleries; same return contract as the reference:
        (scores ndarray [Q, V] float32, txt_ids, vis_ids).  Video embeddings stay on the device (the reference parks
        them on the host and re-uploads them per text batch, model/model.py:1047, :1066).  Galleries above 5e4 videos
        take the predict_batch branch like the reference's (model/model.py:1020-1021)."""
        if _loader_length(vis_loader) > self.LARGE_GALLERY:
            return self.predict_batch(txt_loader, vis_loader, measure, record_emb)
        scores, txt_ids, vis_ids = self.predict_device(txt_loader, vis_loader, measure, record_emb)
        return scores.cpu().numpy(), txt_ids, vis_ids

    LARGE_GALLERY = 5e4   # model/model.py:1020

    def predict_batch(self, txt_loader, vis_loader, measure, record_emb=False):
        """The large-gallery branch (model/model.py:1081-1128).  The reference re-encodes the whole gallery for every text
        batch and fills a dense host matrix -- 40 GB for 10 k x 1 M.  Here the gallery is fused once into a resident
        16-bit index (column j = dataset index j, model/model.py:1118), the queries are fused batch by batch, and the
        first element of the returned triple is a retrieval.RankedScores: the object laff_b200.predictor ranks,
        evaluates and writes the result files from with the fused similarity sweep (no Q x V matrix); np.asarray() of
        it still yields the dense matrix when that is small enough to exist.  record_emb keeps the index for the next
        call, like the reference's video_all_embs cache of predict()."""
        from .retrieval import GalleryIndex, RankedScores
        self.eval()
        if measure != "cosine":
            self.compute_sim(None, None, measure)
        dt = _loss.operand_dtype()
        H = self.opt.multi_head_attention["heads"]
        with torch.no_grad():
            index = getattr(self, "_gallery_index", None) if record_emb else None
            if index is None:
                V = _loader_length(vis_loader)
                g16, self.vis_ids, seen = None, [None] * V, 0
                for output_dict in vis_loader:
                    e16 = self._encode_vis(output_dict, dt)[1]
                    if g16 is None:
                        g16 = torch.empty((V, e16.shape[1] * e16.shape[2]), dtype=e16.dtype, device=e16.device)
                    idxs = torch.as_tensor(np.asarray(output_dict["idxs"]), device=e16.device).long()
                    g16[idxs] = e16.reshape(e16.shape[0], -1)
                    for j, v in zip(np.asarray(output_dict["idxs"]).tolist(), output_dict["vis_ids"]):
                        self.vis_ids[j] = v
                    seen += e16.shape[0]
                if seen != V:
                    raise ops.LaffError("predict_batch: the video loader yielded %d of %d videos" % (seen, V))
                index = GalleryIndex(g16, V, H)
                self._gallery_index = index if record_emb else None
            txt_ids, q = [], []
            for caption_feat_dict, txt_idxs, batch_txt_ids in txt_loader:
                q.append(self.txt_net.encode(caption_feat_dict, out16_dtype=index.g16.dtype)[1])
                txt_ids.extend(batch_txt_ids)
            q16 = torch.cat(q, 0).reshape(len(txt_ids), -1)
        return RankedScores(index, q16), txt_ids, list(self.vis_ids)

    def predict_device(self, txt_loader, vis_loader, measure, record_emb=False):
        """predict() with the score matrix left on the device (CUDA fp32 [Q, V]) for laff_b200.predictor, which ranks,
        evaluates and extracts the written lists there instead of argsorting on the host."""
        self.eval()
        if measure != "cosine":
            self.compute_sim(None, None, measure)
        precision = _loss.get_precision()
        dt = _op_dtype(precision)
        H = self.opt.multi_head_attention["heads"]
        with torch.no_grad():
            if not record_emb or getattr(self, "video_all_embs", None) is None:
                embs, self.video_idxs_list, self.vis_ids = [], [], []
                for output_dict in vis_loader:
                    self.video_idxs_list.append(output_dict["idxs"])
                    embs.append(self._encode_vis(output_dict)[0])
                    self.vis_ids.extend(output_dict["vis_ids"])
                self.video_all_embs = torch.cat(embs, 0)
            V = self.video_all_embs.shape[0]
            order = torch.as_tensor(np.concatenate([np.asarray(i) for i in self.video_idxs_list]), device=self.video_all_embs.device)
            gal = torch.empty_like(self.video_all_embs)
            gal[order] = self.video_all_embs  # column j of the score matrix = dataset index j (model/model.py:1071)
            _, g16 = _loss._operands(gal[:1].reshape(1, -1), gal.reshape(V, -1), H, precision)
            txt_ids, rows = [], []
            for caption_feat_dict, txt_idxs, batch_txt_ids in txt_loader:
                t = self.txt_net(caption_feat_dict)
                q16, _ = _loss._operands(t.reshape(t.shape[0], -1), gal[:1].reshape(1, -1), H, precision)
                rows.append(ops.sim_dense(q16, g16, 1.0 / H))
                txt_ids.extend(batch_txt_ids)
            scores = torch.cat(rows, 0)
        return scores, txt_ids, self.vis_ids


class W2VVPP_MultiHeadAttention(W2VVPP):
    """'LAFF' (model/model.py:1930-2048)."""
    _adam_eps = 1e-4  # model/model.py:2022

    def _init_vis_net(self, opt):
        self.vis_net = VisMutiTransformNetAddAttnetion(opt, opt.vis_fc_layers[0])

    def compute_loss(self, vis_embs, txt_embs, vis_embs_multi_labels=0, txt_embs_multi_labels=0, labels_embs=0):
        if getattr(self.opt, "multi_space", True) and vis_embs.dim() == txt_embs.dim() == 3:
            loss = self.criterion(txt_embs, vis_embs)
        else:
            scores = self.get_txt2vis_matrix(txt_embs, vis_embs, self.opt.measure)  # rows = sentences (model/model.py:2041)
            loss = self.criterion_with_score(scores)
        return loss, {"triplet_loss": loss}

    def change_raw_global_emb_weight(self):
        """Linear decay of the mean-pool residual weight per epoch (model/model.py:1919-1946)."""
        for net, rate in ((self.txt_net, self.opt.txt_attention_global_decay_rate),
                          (self.vis_net, self.opt.vis_attention_global_decay_rate)):
            att = getattr(net, "attention_layer", None)
            if att is not None and hasattr(att, "get_raw_global_emb_weight"):
                att.change_raw_global_emb_weight(max(0.0, rate - 1 + att.get_raw_global_emb_weight()))


class W2VVPP_MutiVisFrameFeat(W2VVPP):
    """'FrameLAFF' / LAFF-ml (model/model.py:2196-2259)."""

    def _init_vis_net(self, opt):
        self.vis_net = VisMutiTransformNetPlusFrameFeat(opt)

    def _encode_vis(self, output_dict, out16_dtype=None):
        return self.vis_net.encode(output_dict["vis_feat_dict"], output_dict["vis_frame_feat_dict"], out16_dtype)


NAME_TO_MODELS = {"LAFF": W2VVPP_MultiHeadAttention, "FrameLAFF": W2VVPP_MutiVisFrameFeat}


def _loader_length(loader) -> int:
    """`vis_loader.dataset.length` as the reference reads it (model/model.py:1020), else len(dataset)."""
    ds = getattr(loader, "dataset", loader)
    n = getattr(ds, "length", None)
    return int(n) if n is not None else len(ds)


def get_model(name, device_, config):
    """model/model.py:2501-2519 for the two LAFF model names."""
    assert name in NAME_TO_MODELS, "%s not supported." % name
    dev = torch.device(device_)
    if dev.type != "cuda":
        raise ops.LaffError("laff_b200 models run on a CUDA (sm_100) device only; got %s" % dev)
    return NAME_TO_MODELS[name](config).float().to(dev)
