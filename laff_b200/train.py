"""One training step of the LAFF model on the device (model/model.py:964-1001).

The reference runs `txt_net` / `vis_net` in train mode under autograd, sums the per-head MarginRankingLoss, calls
`loss.backward()`, `clip_grad_norm_(params, grad_clip)` and `optimizer.step()` (RMSprop lr 1e-4 in the shipped configs).
Here the same step is an explicit forward / backward over the C ABI — no autograd graph:

  forward   per projected feature: tcgen05 GEMM with fused bias + activation (`laff_project`), then dropout + train-mode
            BatchNorm (`laff_transform_train_forward`); no-transform features: tile + train-mode BatchNorm; LAFF pooling
            (`laff_attention_pool`); loss + its gradient w.r.t. both embeddings (`laff_mrl_forward_backward`).
  backward  `laff_attention_pool_backward` -> `laff_transform_train_backward` -> weight gradients dW = dZ^T x as a
            tcgen05 GEMM over K = batch (`laff_transpose_16` operands, `laff_sim_dense`), written straight into `.grad`.
  update    `laff_optimizer_step`: gradient-norm clipping + RMSprop / Adam for every tensor in one pass, no host sync.

Operand precision of the GEMMs: '3-term bf16 split' by default (fp32-grade products — the reference trains in fp32
unless `config.float16`), or plain bf16 / fp16 via `precision=`.  Dropout uses a counter-based mask (seed per step and
feature), not torch's generator: with p > 0 the step is statistically, not bitwise, the reference's.
Restrictions (raise NotImplementedError): the score-matrix loss branch (`multi_space = False`) and the negation-aware
branch.  The GRU sentence encoder trains with the model (laff_b200.text.gru_encode_train / gru_backward).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _capi, ops
from ._capi import LaffError, OptTensor


def _operands(x: torch.Tensor, precision: str, side: int):
    if precision == "bf16x3":
        return ops.split3_16(x, side, torch.bfloat16)
    return ops.cast_pad_16(x, torch.bfloat16 if precision == "bf16" else torch.float16)


def _grad_buffer(p: nn.Parameter) -> torch.Tensor:
    if p.grad is None or p.grad.shape != p.shape or p.grad.device != p.device or not p.grad.is_contiguous():
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


class FusionTrainStep:
    """Train-mode forward + backward of one fusion net: a list of (x, TransformNet) features and its attention layer."""

    def __init__(self, attention, precision: str = "bf16x3"):
        self.att = attention
        # mean-residual / product variants (ablations; shipped: off): omega is read with .item() like the reference does
        # (model/Attention.py:96) — a host synchronisation, so steps of these variants are not graph-captured
        self.capturable = not (getattr(attention, "with_ave", False) or getattr(attention, "mul", False))
        self.precision = precision
        self.cache: Optional[dict] = None

    def forward(self, features: Sequence, seed: int, seed_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
        """features: [(x fp32 on the device, TransformNet)].  seed (+ the device step counter seed_dev) keys the dropout
        masks."""
        att = self.att
        H, dh = att.multi_heads, att.dim_per_head
        D = H * dh
        # Every feature's chain (operand casts -> projection GEMM -> dropout / BatchNorm) is independent of the others
        # and, at B = 128, fills a fraction of the SMs: each runs on its own stream, forked from and joined into the
        # current one (inside a graph capture: parallel branches).
        dev = features[0][0].device
        cur = torch.cuda.current_stream(dev)
        lanes = self._lanes(dev, len(features))
        items = []
        for i, f in enumerate(features):
            x, tn = f[0], f[1]
            p = float(tn.dropout_p or 0.0)
            it = {"x": x, "tn": tn, "p": p, "want_dx": len(f) > 2 and bool(f[2])}
            lanes[i].wait_stream(cur)
            with torch.cuda.stream(lanes[i]):
                if tn.fc1 is not None:
                    w16 = _operands(tn.fc1.weight.detach(), self.precision, 1)
                    a = ops.project(_operands(x, self.precision, 0), w16, tn.fc1.bias.detach(), tn.activation_name)
                    it["a"] = a
                    if p > 0 or tn.bn1 is not None:
                        y, mask, sm, si = ops.transform_train_forward(a, D, p, seed * 131 + i, tn.bn1, seed_dev=seed_dev)
                    else:
                        y, mask, sm, si = a, None, None, None
                else:
                    y, mask, sm, si = ops.transform_train_forward(x, D, p, seed * 131 + i, tn.bn1, seed_dev=seed_dev)
                if tn.bn1 is not None:
                    tn.bn1.num_batches_tracked += 1
            it.update(y=y, mask=mask, sm=sm, si=si)
            items.append(it)
        for i in range(len(features)):
            cur.wait_stream(lanes[i])
        ps = [att.attention_layer[h].embedding_common[0] for h in range(H)]
        w = torch.cat([q.weight.detach().view(1, -1) for q in ps], 0).float().contiguous()
        b = torch.cat([q.bias.detach().view(1) for q in ps], 0).float().contiguous()
        omega = float(att.attention_layer[0].global_emb_weight_net.weight.item()) if att.with_ave else 0.0
        out, _, _ = ops.attention_pool([{"y": it["y"]} for it in items], w, b, H, dh, att.with_ave, att.mul, omega=omega)
        self.cache = {"items": items, "w": w, "b": b, "heads": ps, "omega": omega}
        return out

    def _lanes(self, dev, n: int):
        """n side streams on `dev`, one per feature chain (kept for the life of the step object)."""
        ls = getattr(self, "_lane_streams", None)
        if ls is None or len(ls) < n or ls[0].device != dev:
            ls = self._lane_streams = [torch.cuda.Stream(dev) for _ in range(n)]
        return ls

    def _feature_backward(self, idx: int, it: dict, dy: torch.Tensor, dxs: Dict[int, torch.Tensor]) -> None:
        """Backward of one feature chain down to its parameters (and its input when that is itself computed)."""
        tn = it["tn"]
        dgamma = _grad_buffer(tn.bn1.weight) if tn.bn1 is not None else None
        dbeta = _grad_buffer(tn.bn1.bias) if tn.bn1 is not None else None
        if tn.fc1 is None:
            if tn.bn1 is not None or it["want_dx"]:
                dzt = ops.transform_train_backward(dy, None, it["x"], it["mask"], it["p"], "none", tn.bn1, it["sm"], it["si"],
                                                   want_dz=it["want_dx"], dgamma=dgamma, dbeta=dbeta)
                if it["want_dx"]:
                    dxs[idx] = ops.fold_tiles(dzt, it["x"].shape[1])
            return
        dbias = _grad_buffer(tn.fc1.bias)
        dz = ops.transform_train_backward(dy, it["a"], None, it["mask"], it["p"], tn.activation_name, tn.bn1, it["sm"],
                                          it["si"], dgamma=dgamma, dbeta=dbeta, dbias=dbias)
        terms = 3 if self.precision == "bf16x3" else 1
        dt = torch.float16 if self.precision == "fp16" else torch.bfloat16
        if it["want_dx"]:  # the input is itself computed (GRU sentence feature): dx = dz @ W
            wT16 = ops.transpose_16(tn.fc1.weight.detach(), dt, terms, 1)
            dz16 = ops.split3_16(dz, 0, torch.bfloat16) if terms == 3 else ops.cast_pad_16(dz, dt)
            dxs[idx] = ops.project(dz16, wT16, None, "none")
        ops.sim_dense(ops.transpose_16(dz, dt, terms, 0), ops.transpose_16(it["x"], dt, terms, 1), 1.0,
                      out=_grad_buffer(tn.fc1.weight))

    def backward(self, dout: torch.Tensor) -> Dict[int, torch.Tensor]:
        """Writes every parameter gradient of the net into `.grad`; returns {feature index: d loss / d x} for the
        no-transform features given as (x, TransformNet, True) — inputs that are themselves computed (LAFF-ml's pooled
        frame feature)."""
        c = self.cache
        if c is None:
            raise LaffError("FusionTrainStep.backward called before forward")
        att = self.att
        H, dh = att.multi_heads, att.dim_per_head
        dev = dout.device
        dout = dout.reshape(dout.shape[0], -1)
        dw = getattr(self, "_dw", None)
        if dw is None or dw.device != dev:
            self._dw = dw = torch.zeros((H, dh), dtype=torch.float32, device=dev)
            self._dc = torch.zeros(H, dtype=torch.float32, device=dev)
        dys = ops.attention_pool_backward([it["y"] for it in c["items"]], c["w"], c["b"], H, dh, dout, dw, self._dc,
                                          with_ave=att.with_ave, mul=att.mul, omega=c["omega"])
        for h, q in enumerate(c["heads"]):  # per-head parameters: gradients are views of the two stacked buffers
            q.weight.grad = dw[h:h + 1]
            q.bias.grad = self._dc[h:h + 1]
        dxs: Dict[int, torch.Tensor] = {}
        cur = torch.cuda.current_stream(dev)
        lanes = self._lanes(dev, len(c["items"]))
        for idx, (it, dy) in enumerate(zip(c["items"], dys)):
            lanes[idx].wait_stream(cur)
            with torch.cuda.stream(lanes[idx]):
                self._feature_backward(idx, it, dy, dxs)
        for idx in range(len(c["items"])):
            cur.wait_stream(lanes[idx])
        self.cache = None
        return dxs


class DeviceGradScaler:
    """torch.cuda.amp.GradScaler as the reference's float16 training branch drives it (model/model.py:793, :970-989),
    with its state -- scale S, growth tracker, found-inf flag, skipped-step count -- in four device words updated by
    `laff_optimizer_step_scaled`, so the step stays free of host synchronisation and graph-capturable.  The gradients
    of this implementation are fp32 and never scaled; S enters where the reference's ordering makes it observable:
    `clip_grad_norm_` runs on the scaled gradients (the effective clip threshold on the true gradients is
    grad_clip / S), and a step whose scaled parameter gradients would overflow fp16 is skipped and halves S."""

    def __init__(self, init_scale: float = 65536.0, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, overflow_limit: float = 65520.0):
        self.init_scale, self.growth_factor, self.backoff_factor = float(init_scale), float(growth_factor), float(backoff_factor)
        self.growth_interval, self.overflow_limit = int(growth_interval), float(overflow_limit)
        self._state = None

    def state(self, device) -> torch.Tensor:
        if self._state is None or self._state.device != device:
            st = torch.zeros(4, dtype=torch.int32, device=device)
            st[:1].view(torch.float32).fill_(self.init_scale)
            self._state = st
        return self._state

    def _host(self):
        if self._state is None:
            return self.init_scale, 0, 0, 0
        h = self._state.cpu()
        return float(h[:1].view(torch.float32)[0]), int(h[1]), int(h[2]), int(h[3])

    def get_scale(self) -> float:
        return self._host()[0]

    def skipped_steps(self) -> int:
        return self._host()[3]

    def last_step_skipped(self) -> bool:
        return bool(self._host()[2])

    def state_dict(self):
        s, tracker, _, _ = self._host()
        return {"scale": s, "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": tracker}

    def load_state_dict(self, sd):
        self.growth_factor, self.backoff_factor = float(sd["growth_factor"]), float(sd["backoff_factor"])
        self.growth_interval = int(sd["growth_interval"])
        self.init_scale = float(sd["scale"])
        if self._state is not None:
            self._state[:1].view(torch.float32).fill_(self.init_scale)
            self._state[1] = int(sd.get("_growth_tracker", 0))


class DeviceOptimizer:
    """clip_grad_norm_ + torch.optim.RMSprop / Adam semantics for a fixed list of parameters, one fused device pass per
    step (`laff_optimizer_step`).  State tensors are allocated on first use; parameters whose `.grad` is None at the
    first step are skipped for good (like parameters that never receive a gradient in the reference's model)."""

    def __init__(self, params: Sequence[nn.Parameter], kind: str = "rmsprop", lr: float = 1e-4, alpha: float = 0.99,
                 betas=(0.9, 0.999), eps: Optional[float] = None, max_grad_norm: float = 0.0, scaler: Optional["DeviceGradScaler"] = None):
        if kind not in ("rmsprop", "adam"):
            raise LaffError("optimizer %r: the reference trains with 'rmsprop' or 'adam'" % kind)
        self.params = [p for p in params]
        self.kind, self.lr, self.alpha, self.betas = kind, float(lr), float(alpha), betas
        self.eps = float(eps) if eps is not None else 1e-8
        self.max_grad_norm = float(max_grad_norm)
        self.step_count = 0
        self._built = None
        self._state = {}
        self.scaler = scaler   # the reference's float16 branch (model/model.py:970-989): see DeviceGradScaler
        self.param_groups = [{"lr": self.lr}]  # lr schedulers of the caller read / write this like torch's

    def _build(self):
        live = [p for p in self.params if p.grad is not None]
        if not live:
            raise LaffError("DeviceOptimizer.step: no parameter has a gradient")
        dev = live[0].device
        # a rebuild (a parameter or gradient was re-allocated) keeps the running averages of the parameters it already
        # tracked, keyed by parameter object, and the device-side step count; only new parameters start from zero
        old = self._state if getattr(self, "_state", None) else {}
        self._state = {}
        for p in live:
            s1, s2 = old.get(id(p), (None, None))
            if s1 is None or s1.shape != p.shape or s1.device != p.device:
                s1 = torch.zeros_like(p, memory_format=torch.contiguous_format)
                s2 = torch.zeros_like(p, memory_format=torch.contiguous_format) if self.kind == "adam" else None
            self._state[id(p)] = (s1, s2)
        self.state1 = [self._state[id(p)][0] for p in live]
        self.state2 = [self._state[id(p)][1] for p in live]
        prev_step = self._built["step_dev"] if self._built is not None else None
        arr = (OptTensor * len(live))()
        for i, p in enumerate(live):
            if not p.is_contiguous() or not p.grad.is_contiguous() or p.dtype != torch.float32:
                raise LaffError("DeviceOptimizer needs contiguous fp32 parameters and gradients")
            arr[i].param, arr[i].grad, arr[i].grad_out = p.data_ptr(), p.grad.data_ptr(), p.grad.data_ptr()
            arr[i].state1 = self.state1[i].data_ptr()
            arr[i].state2 = self.state2[i].data_ptr() if self.state2[i] is not None else None
            arr[i].n = p.numel()
        sizes = (C.c_longlong * len(live))(*[p.numel() for p in live])
        lib = _capi.lib()
        n_blocks = lib.laff_optimizer_blocks(sizes, len(live), None, None, 0)
        if n_blocks < 0:
            _capi.check(n_blocks, "laff_optimizer_blocks")
        bt = (C.c_int * n_blocks)()
        bs = (C.c_longlong * n_blocks)()
        lib.laff_optimizer_blocks(sizes, len(live), bt, bs, n_blocks)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self._built = {
            "live": live, "ptrs": [(p.data_ptr(), p.grad.data_ptr()) for p in live], "desc": raw,
            "bt": torch.tensor(list(bt), dtype=torch.int32, device=dev), "bs": torch.tensor(list(bs), dtype=torch.int64, device=dev),
            "partial": torch.empty(n_blocks, dtype=torch.float64, device=dev),
            "partial_max": torch.empty(n_blocks, dtype=torch.float32, device=dev),
            "ctl": torch.zeros(2, dtype=torch.float32, device=dev),
            "norm": torch.zeros(1, dtype=torch.float64, device=dev), "n_blocks": n_blocks,
            # the device word is the authoritative step count (graph replays advance it, not the Python counter)
            "step_dev": prev_step if prev_step is not None else torch.full((1,), self.step_count, dtype=torch.int64, device=dev),
            "lr_dev": torch.zeros(1, dtype=torch.float32, device=dev), "lr_host": None}

    def step(self) -> torch.Tensor:
        """Returns the (device) total gradient norm before clipping.  The step count and the learning rate are read
        from device words, so the call can sit inside a captured CUDA graph; `param_groups[0]['lr']` is pushed to the
        device whenever it changes (outside any capture)."""
        if self._built is None:
            self._build()
        b = self._built
        if [(p.data_ptr(), p.grad.data_ptr()) for p in b["live"]] != b["ptrs"]:
            self._build()  # a parameter or gradient was re-allocated
            b = self._built
        self.sync_lr()
        self.step_count += 1
        first = self.alpha if self.kind == "rmsprop" else self.betas[0]
        if self.scaler is not None:
            sc = self.scaler
            _capi.call("laff_optimizer_step_scaled", ops._ptr(b["desc"]), ops._ptr(b["bt"]), ops._ptr(b["bs"]), b["n_blocks"],
                       0 if self.kind == "rmsprop" else 1, float(first), float(self.betas[1]), self.eps, self.max_grad_norm,
                       ops._ptr(b["partial"]), ops._ptr(b["partial_max"]), ops._ptr(b["norm"]), ops._ptr(b["step_dev"]),
                       ops._ptr(b["lr_dev"]), ops._ptr(sc.state(b["desc"].device)), sc.growth_factor, sc.backoff_factor,
                       sc.growth_interval, sc.overflow_limit, ops._ptr(b["ctl"]), ops._stream(b["desc"]))
            return b["norm"]
        b["step_dev"].add_(1)
        _capi.call("laff_optimizer_step", ops._ptr(b["desc"]), ops._ptr(b["bt"]), ops._ptr(b["bs"]), b["n_blocks"],
                   0 if self.kind == "rmsprop" else 1, float(b["lr_host"]), float(first), float(self.betas[1]), self.eps,
                   max(1, self.step_count), self.max_grad_norm, ops._ptr(b["partial"]), ops._ptr(b["norm"]),
                   ops._ptr(b["step_dev"]), ops._ptr(b["lr_dev"]), ops._stream(b["desc"]))
        return b["norm"]

    def sync_lr(self) -> None:
        b = self._built
        if b is None:
            return
        lr = float(self.param_groups[0]["lr"])
        if b.get("lr_host") != lr:
            if torch.cuda.is_current_stream_capturing():
                raise LaffError("the learning rate changed inside a CUDA graph capture")
            b["lr_dev"].fill_(lr)
            b["lr_host"] = lr

    def zero_grad(self):
        pass  # every gradient buffer is overwritten by the next backward
