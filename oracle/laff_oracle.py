"""CPU oracle for the LAFF retrieval hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy restatement of the reference's algorithm (ruc-aimc-lab/LAFF), function by function, each citing the
reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl
reference`` legs of ``bench.py`` may import this module; the product path (``laff_b200``) never does and fails loudly
when its CUDA library is missing.

Pinning: the reference ships no golden vectors or known-answer tests, so this oracle is pinned against
outputs of the reference itself, generated in the build container by ``tests/golden/make_golden.py`` (which imports
the unmodified reference from /root/reference) and committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function here against them.

The arithmetic lives in third-party code the reference calls (PyTorch/ATen ``nn.Linear``, ``mm``, ``softmax``,
``BatchNorm1d``; numpy ``argsort``/``median``; reference pins torch 1.7.1 / numpy 1.20.1, requirements.txt:1-5); the
formulas below are those libraries' published semantics in float32 (pass ``dtype=np.float64`` for a high-precision
variant used to bound rounding noise).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------------------------
# normalisation and similarity
# ----------------------------------------------------------------------------------------------------------------
def l2norm(X: np.ndarray, eps: float = 1e-13, axis: int = 1) -> np.ndarray:
    """loss.py:8-13 — X / (sqrt(sum(X^2)) + eps + 1e-14); the additions happen in X's dtype like torch does."""
    dt = X.dtype.type
    norm = np.sqrt(np.sum(X * X, axis=axis, keepdims=True, dtype=X.dtype))
    norm = norm + dt(eps)
    norm = norm + dt(1e-14)
    return X / norm


def l2norm_np(X: np.ndarray) -> np.ndarray:
    """evaluation.py:11-16 — X / (||X|| + 1e-10)."""
    norm = np.linalg.norm(X, axis=1, keepdims=True)
    return 1.0 * X / (norm + 1e-10)


def cosine_sim(query: np.ndarray, retrio: np.ndarray) -> np.ndarray:
    """loss.py:30-34 — l2norm both sides, then query @ retrio^T."""
    return l2norm(query) @ l2norm(retrio).T


def cosine_sim_np(query_embs: np.ndarray, retro_embs: np.ndarray) -> np.ndarray:
    """evaluation.py:44-50."""
    return l2norm_np(query_embs).dot(l2norm_np(retro_embs).T)


def compute_sim(query_embs, retro_embs, measure: str = "cosine"):
    """model/model.py:1567-1578 / evaluation.py:53-61 — only 'cosine' is used by the shipped configs."""
    if measure == "cosine":
        return cosine_sim(query_embs, retro_embs)
    if measure in ("hist", "euclidean"):
        raise Exception("Not implemented")
    raise Exception("%s is invalid" % measure)


def txt2vis_matrix(txt_embs: np.ndarray, vis_embs: np.ndarray) -> np.ndarray:
    """model/model.py:1003-1016 — 2-D: cosine_sim; 3-D [N, H, d]: mean over heads of the per-head cosine_sim."""
    if txt_embs.ndim == 2 and vis_embs.ndim == 2:
        return cosine_sim(txt_embs, vis_embs)
    H = vis_embs.shape[1]
    sims = [cosine_sim(np.ascontiguousarray(txt_embs[:, h, :]), np.ascontiguousarray(vis_embs[:, h, :])) for h in range(H)]
    return np.mean(np.stack(sims, 0), axis=0, dtype=sims[0].dtype)


def mm_mean_heads(q_hat: np.ndarray, g_hat: np.ndarray, heads: int) -> np.ndarray:
    """The `query.mm(retrio.t())` (loss.py:34) + mean over heads (model/model.py:1014) stage alone, on operands that
    are already normalised and rounded to the tensor-core dtype: the stage the similarity GEMM replaces.  float64."""
    return (q_hat.astype(np.float64) @ g_hat.astype(np.float64).T) / float(heads)


# ----------------------------------------------------------------------------------------------------------------
# projection + LAFF block
# ----------------------------------------------------------------------------------------------------------------
def activation_fn(x: np.ndarray, name) -> np.ndarray:
    """model/model.py:234-241."""
    if name == "tanh":
        return np.tanh(x)
    if name == "relu":
        return np.maximum(x, 0)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    return x


def batchnorm_eval(x, weight, bias, mean, var, eps=1e-5):
    """nn.BatchNorm1d in eval mode (model/model.py:232, :273-274): (x - mean) / sqrt(var + eps) * weight + bias."""
    dt = x.dtype.type
    return (x - mean) / np.sqrt(var + dt(eps)) * weight + bias


def transform_net(x: np.ndarray, sd: Mapping[str, np.ndarray], prefix: str, activation="tanh") -> np.ndarray:
    """TransformNet.forward, eval mode (model/model.py:257-276): FC -> activation -> (dropout = id) -> BN.
    A stage exists iff its parameters are in the state dict (fc1.* / bn1.*)."""
    y = x
    if prefix + "fc1.weight" in sd:
        y = y @ sd[prefix + "fc1.weight"].T + sd[prefix + "fc1.bias"]
        y = activation_fn(y, activation)
    if prefix + "bn1.weight" in sd:
        y = batchnorm_eval(y, sd[prefix + "bn1.weight"], sd[prefix + "bn1.bias"], sd[prefix + "bn1.running_mean"],
                           sd[prefix + "bn1.running_var"])
    return y


def softmax(x: np.ndarray, axis: int) -> np.ndarray:
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True, dtype=x.dtype)


def attention_1(local_embs: np.ndarray, w: np.ndarray, c, with_ave: bool, mul: bool, omega: float = 1.0):
    """Attention_1.forward (model/Attention.py:78-105). local_embs [B, L, d]; w [d]; c scalar.
    Returns (new_global [B, d] unit norm, weights [B, L] as stored in self.weights)."""
    dt = local_embs.dtype.type
    raw_global = np.mean(local_embs, axis=1, dtype=local_embs.dtype)              # :81
    common = local_embs
    if mul:
        common = local_embs * raw_global[:, None, :]                              # :83-86
    logits = common @ w + dt(c)                                                   # :88
    weights = softmax(logits, axis=1)                                             # :89
    new_global = weights[:, :, None] * local_embs                                 # :93
    stored = weights
    if with_ave:
        stored = weights + dt(omega) * dt(1.0) / dt(weights.shape[1])             # :97
        new_global = new_global + dt(omega) * raw_global[:, None, :]              # :99
    new_global = np.sum(new_global, axis=1, dtype=local_embs.dtype)               # :101
    return l2norm(new_global, eps=0), stored                                      # :103


def multi_head_attention(local_embs: np.ndarray, sd: Mapping[str, np.ndarray], prefix: str, heads: int,
                         with_ave: bool, mul: bool):
    """Multi_head_MyApply_Attention.forward with split_head=True (model/Attention.py:508-531).
    local_embs [B, L, D] -> ([B, H, D/H], weights [B, H, L])."""
    B, L, D = local_embs.shape
    dh = D // heads
    x = local_embs.reshape(B, L, heads, dh)                                       # :517
    outs, atts = [], []
    for h in range(heads):                                                        # :525-527
        p = "%sattention_layer.%d." % (prefix, h)
        w = sd[p + "embedding_common.0.weight"].reshape(-1)
        c = sd[p + "embedding_common.0.bias"].reshape(-1)[0]
        omega = float(sd[p + "global_emb_weight_net.weight"].reshape(-1)[0])
        o, a = attention_1(np.ascontiguousarray(x[:, :, h, :]), w, c, with_ave, mul, omega)
        outs.append(o)
        atts.append(a)
    return np.stack(outs, 1), np.stack(atts, 1)                                   # :529


def vis_net_forward(vis_input: Mapping[str, np.ndarray], sd: Mapping[str, np.ndarray], no_transform: Sequence[str],
                    heads: int, with_ave=False, mul=False, activation="tanh"):
    """VisMutiTransformNetAddAttnetion.forward -> VisMutiTransformNet.forward, eval (model/model.py:1858-1876,
    :1807-1827): per-feature TransformNet (no-transform features tiled x heads, :1822-1823), stack in dict order
    (:1862), multi-head LAFF."""
    feats = []
    for name, x in vis_input.items():
        if name in no_transform:
            x = np.tile(x, (1, heads))
        feats.append(transform_net(x, sd, "VisMutiTransformNet.%s." % name, activation))
    return multi_head_attention(np.stack(feats, 1), sd, "attention_layer.", heads, with_ave, mul)


TXT_ENCODER_ORDER = ("rnn_encoder", "bow_encoder", "w2v_encoder", "CLIP_encoder")  # model/model.py:573-620
TXT_FEATURE_KEY = {"rnn_encoder": "gru", "bow_encoder": "bow", "w2v_encoder": "w2v", "CLIP_encoder": "clip"}


def txt_net_forward(txt_feats: Mapping[str, np.ndarray], sd: Mapping[str, np.ndarray], no_transform: Sequence[str],
                    heads: int, with_ave=False, mul=False, activation="tanh"):
    """MultiScaleTxtEncoderAttention.forward, eval (model/model.py:1663-1705) with the encoder outputs given as
    features {'gru','bow','w2v','clip'}: tile no-transform (:1675-1676), per-encoder TransformNet (:1678), stack in
    encoder_name_list order (:1683), multi-head LAFF (:1703)."""
    feats = []
    for enc in TXT_ENCODER_ORDER:
        key = TXT_FEATURE_KEY[enc]
        if key not in txt_feats:
            continue
        x = txt_feats[key]
        if enc in no_transform:
            x = np.tile(x, (1, heads))
        feats.append(transform_net(x, sd, "transform_layer.%s_transform." % enc, activation))
    return multi_head_attention(np.stack(feats, 1), sd, "attention_layer.", heads, with_ave, mul)


def frame_attention_forward(frames: np.ndarray, sd: Mapping[str, np.ndarray], frame_feat: str):
    """The frame-level stage of VisMutiTransformNetPlusFrameFeat.forward (model/model.py:2160-2173): Attention_1(dim,
    with_ave=False, mul=False) over the frame axis of each video; the mask slice acts on the batch axis and is a no-op,
    so zero-padded frames take part in the softmax.  frames [B, F, dim] -> [B, dim]."""
    p = "frame_attention.%s.0." % frame_feat
    w = sd[p + "embedding_common.0.weight"].reshape(-1)
    c = sd[p + "embedding_common.0.bias"].reshape(-1)[0]
    out, _ = attention_1(frames, w, c, with_ave=False, mul=False)
    return out


def frame_vis_net_forward(vis_input: Mapping[str, np.ndarray], frames: np.ndarray, frame_feat: str,
                          sd: Mapping[str, np.ndarray], no_transform: Sequence[str], heads: int, activation="tanh"):
    """VisMutiTransformNetPlusFrameFeat.forward, eval (model/model.py:2147-2190): the pooled frame feature joins the
    video-level features (appended last), no-transform ones are tiled (:2182-2184), TransformNet each, stack, LAFF."""
    feats_in = dict(vis_input)
    feats_in[frame_feat] = frame_attention_forward(frames, sd, frame_feat)
    feats = []
    for name, x in feats_in.items():
        if name in no_transform:
            x = np.tile(x, (1, heads))
        feats.append(transform_net(x, sd, "%s." % name, activation))
    return multi_head_attention(np.stack(feats, 1), sd, "vis_attention_layer.", heads, False, False)


# ----------------------------------------------------------------------------------------------------------------
# ranking and metrics
# ----------------------------------------------------------------------------------------------------------------
def argsort_rank(scores: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """predictor.py:232-244 / evaluation.py:71-79 — inds = np.argsort(row)[::-1]; rank = position of the ground truth.
    numpy's default sort is not stable, so exact ties are resolved however numpy resolves them."""
    inds = np.argsort(scores, axis=1)
    return np.array([np.where(inds[i][::-1] == gt[i])[0][0] for i in range(scores.shape[0])], dtype=np.int64)


def tie_rule_rank(scores: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """The documented convention = np.argsort(kind='stable')[::-1]: order by (score desc, index desc):
    rank0 = #{j != gt: s_j > s_gt} + #{j > gt: s_j == s_gt}."""
    Q, V = scores.shape
    sg = scores[np.arange(Q), gt][:, None]
    cols = np.arange(V)[None, :]
    beats = (scores > sg) | ((scores == sg) & (cols > gt[:, None]))
    beats &= cols != gt[:, None]
    return beats.sum(1).astype(np.int64)


def rank_bounds(scores: np.ndarray, gt: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Every rank an argsort-based search can return: #{s_j > s_gt} <= rank0 <= #{s_j > s_gt} + #{j != gt: s_j == s_gt}.
    numpy's default argsort (predictor.py:232) is not stable (introsort / SIMD quicksort depending on the build), so
    on exact ties the reference's own answer is only pinned to this interval."""
    Q, V = scores.shape
    sg = scores[np.arange(Q), gt][:, None]
    lo = (scores > sg).sum(1)
    eq = (scores == sg).sum(1) - 1
    return lo.astype(np.int64), (lo + eq).astype(np.int64)


def tie_rule_topk(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k under the same order: np.argsort(kind='stable')[::-1][:k]."""
    inds = np.argsort(scores, axis=1, kind="stable")[:, ::-1][:, :k]
    return np.take_along_axis(scores, inds, 1), inds.astype(np.int64)


def metrics_from_rank0(rank0: np.ndarray) -> Tuple[float, float, float, float, float, float]:
    """evaluation.py:81-89 — from 0-based ranks: R@1/5/10 (%), MedR = floor(median)+1, MeanR = mean+1, MIR."""
    ranks = np.asarray(rank0, dtype=np.float64)
    n = len(ranks)
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / n
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / n
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / n
    medr = np.floor(np.median(ranks)) + 1
    meanr = ranks.mean() + 1
    mir = (1.0 / (ranks + 1)).mean()
    return (r1, r5, r10, medr, meanr, mir)


def eval_qry2retro(qry2retro_sim: np.ndarray, n_qry: int = 1):
    """evaluation.py:64-89 — ground truth of row i is column i / n_qry."""
    assert qry2retro_sim.shape[0] / qry2retro_sim.shape[1] == n_qry, qry2retro_sim.shape
    gt = (np.arange(qry2retro_sim.shape[0]) / n_qry).astype(np.int64)
    return metrics_from_rank0(argsort_rank(qry2retro_sim, gt))


def eval_label_matrix(label_matrix: np.ndarray):
    """evaluation.py:92-109 — label_matrix[i, p] = 1 iff the item at sorted position p is a ground truth of query i.
    Returns (r1, r5, r10, medr, meanr, mir, mAP) from 1-based ranks."""
    label_matrix = label_matrix.astype(int)
    n = label_matrix.shape[0]
    ranks = np.zeros(n)
    aps = np.zeros(n)
    for i in range(n):
        pos = np.where(label_matrix[i] == 1)[0] + 1
        ranks[i] = pos[0]
        aps[i] = np.mean([(j + 1.0) / pos[j] for j in range(len(pos))])
    r1, r5, r10 = [100.0 * np.mean([x <= k for x in ranks]) for k in (1, 5, 10)]
    return (r1, r5, r10, np.floor(np.median(ranks)), ranks.mean(), (1.0 / ranks).mean(), aps.mean())


# ----------------------------------------------------------------------------------------------------------------
# training step: explicit forward / backward of what the reference leaves to autograd
# ----------------------------------------------------------------------------------------------------------------
def _act_grad(a: np.ndarray, name) -> np.ndarray:
    if name == "tanh":
        return 1.0 - a * a
    if name == "sigmoid":
        return a * (1.0 - a)
    if name == "relu":
        return (a > 0).astype(a.dtype)
    return np.ones_like(a)


def fusion_train_forward(feats: Sequence[Tuple[str, np.ndarray]], sd: dict, prefixes: Mapping[str, str], att_prefix: str,
                         heads: int, no_transform: Sequence[str], activation="tanh", momentum=0.1, eps=1e-5, with_ave=False,
                         mul=False):
    """Train-mode forward of one fusion net (TransformNet.forward with batch-statistics BatchNorm, dropout p = 0;
    model/model.py:257-276, :1807-1876) followed by the multi-head LAFF block.  Updates the running statistics in `sd`
    in place like nn.BatchNorm1d.  Returns (embeddings [B, H, d_h], cache for fusion_train_backward)."""
    items = []
    for name, x in feats:
        pre = prefixes[name]
        it = {"name": name, "pre": pre, "x": x.astype(np.float64)}
        if name in no_transform:
            y = np.tile(it["x"], (1, heads))
        else:
            z = it["x"] @ sd[pre + "fc1.weight"].astype(np.float64).T + sd[pre + "fc1.bias"]
            y = activation_fn(z, activation)
            it["a"] = y
        if pre + "bn1.weight" in sd:
            B = y.shape[0]
            mean, var = y.mean(0), y.var(0)
            it["xhat"] = (y - mean) / np.sqrt(var + eps)
            it["invstd"] = 1.0 / np.sqrt(var + eps)
            y = it["xhat"] * sd[pre + "bn1.weight"] + sd[pre + "bn1.bias"]
            sd[pre + "bn1.running_mean"] = ((1 - momentum) * sd[pre + "bn1.running_mean"] + momentum * mean).astype(np.float32)
            sd[pre + "bn1.running_var"] = ((1 - momentum) * sd[pre + "bn1.running_var"] + momentum * var * B / (B - 1)).astype(np.float32)
            if pre + "bn1.num_batches_tracked" in sd:
                sd[pre + "bn1.num_batches_tracked"] = sd[pre + "bn1.num_batches_tracked"] + 1
        it["y"] = y
        items.append(it)
    Y = np.stack([it["y"] for it in items], 1)                                  # [B, L, D]
    B, L, D = Y.shape
    dh = D // heads
    Yh = Y.reshape(B, L, heads, dh)
    W = np.stack([sd["%sattention_layer.%d.embedding_common.0.weight" % (att_prefix, h)].reshape(-1) for h in range(heads)]).astype(np.float64)
    c = np.array([sd["%sattention_layer.%d.embedding_common.0.bias" % (att_prefix, h)].reshape(-1)[0] for h in range(heads)], dtype=np.float64)
    r = Yh.mean(1)                                                                # [B, H, dh]   model/Attention.py:81
    common = Yh * r[:, None] if mul else Yh                                       # :83-86
    e = np.einsum("blhd,hd->bhl", common, W) + c[None, :, None]
    p = softmax(e, axis=2)
    omega = np.array([float(sd["%sattention_layer.%d.global_emb_weight_net.weight" % (att_prefix, h)].reshape(-1)[0])
                      for h in range(heads)]) if with_ave else np.zeros(heads)    # read with .item(): a constant (:96)
    g = np.einsum("bhl,blhd->bhd", p + omega[None, :, None], Yh)
    nrm = np.sqrt((g * g).sum(-1, keepdims=True))
    out = g / (nrm + 1e-14)
    return out, {"items": items, "Yh": Yh, "W": W, "p": p, "g": g, "nrm": nrm, "out": out, "heads": heads, "att_prefix": att_prefix,
                 "activation": activation, "r": r, "common": common, "omega": omega, "mul": mul}


def fusion_train_backward(cache: dict, dout: np.ndarray, sd: Mapping[str, np.ndarray]) -> dict:
    """Gradients of every parameter of one fusion net given d loss / d embeddings [B, H, d_h]."""
    Yh, W, p, g, nrm, out, heads = (cache[k] for k in ("Yh", "W", "p", "g", "nrm", "out", "heads"))
    B, L, H, dh = Yh.shape
    inv = 1.0 / (nrm + 1e-14)
    dot = (out * dout).sum(-1, keepdims=True)
    dg = (dout - out * dot * nrm * inv) * inv
    dp = np.einsum("bhd,blhd->bhl", dg, Yh)
    de = p * (dp - (p * dp).sum(-1, keepdims=True))
    dY = np.einsum("bhl,bhd->blhd", p + cache["omega"][None, :, None], dg)
    dcommon = np.einsum("bhl,hd->blhd", de, W)
    if cache["mul"]:
        dY = dY + dcommon * cache["r"][:, None] + (dcommon * Yh).sum(1, keepdims=True) / L
    else:
        dY = dY + dcommon
    grads = {}
    for h in range(H):
        pre = "%sattention_layer.%d.embedding_common.0." % (cache["att_prefix"], h)
        grads[pre + "weight"] = np.einsum("bl,bld->d", de[:, h, :], cache["common"][:, :, h, :]).reshape(1, dh)
        grads[pre + "bias"] = de[:, h, :].sum().reshape(1)
    dY = dY.reshape(B, L, H * dh)
    cache["dx"] = {}
    for l, it in enumerate(cache["items"]):
        pre, d = it["pre"], dY[:, l, :]
        if "xhat" in it:
            grads[pre + "bn1.weight"] = (d * it["xhat"]).sum(0)
            grads[pre + "bn1.bias"] = d.sum(0)
            gam = sd[pre + "bn1.weight"]
            d = gam * it["invstd"] * (d - d.mean(0) - it["xhat"] * (d * it["xhat"]).mean(0))
        if "a" in it:
            dz = d * _act_grad(it["a"], cache["activation"])
            grads[pre + "fc1.weight"] = dz.T @ it["x"]
            grads[pre + "fc1.bias"] = dz.sum(0)
            cache["dx"][it["name"]] = dz @ sd[pre + "fc1.weight"].astype(np.float64)   # needed when x is computed (GRU feature)
        else:  # tiled feature: x.repeat(1, heads) backward (needed when x is itself computed, LAFF-ml)
            cache["dx"][it["name"]] = d.reshape(B, H, -1).sum(1)
    return grads


def frame_attention_train(frames: np.ndarray, w: np.ndarray, dout: Optional[np.ndarray] = None):
    """Frame-level Attention_1 (with_ave = mul = False) per video over all F (zero-padded) frames
    (model/model.py:2167-2173).  Without dout: pooled features [B, dim].  With dout: (d w [dim], d c scalar)."""
    x = frames.astype(np.float64)
    e = x @ w.astype(np.float64)                                        # the bias cancels in the softmax
    p = softmax(e, axis=1)
    g = np.einsum("bf,bfd->bd", p, x)
    nrm = np.sqrt((g * g).sum(-1, keepdims=True))
    out = g / (nrm + 1e-14)
    if dout is None:
        return out
    inv = 1.0 / (nrm + 1e-14)
    dg = (dout - out * (out * dout).sum(-1, keepdims=True) * nrm * inv) * inv
    dp = np.einsum("bd,bfd->bf", dg, x)
    de = p * (dp - (p * dp).sum(-1, keepdims=True))
    return np.einsum("bf,bfd->d", de, x), de.sum()


FP16_OVERFLOW = 65520.0   # smallest magnitude that rounds to inf in IEEE half precision


def clip_and_step(sd: dict, grads: Mapping[str, np.ndarray], state: dict, optimizer="rmsprop", lr=1e-4, max_norm=2.0,
                  alpha=0.99, betas=(0.9, 0.999), eps=None, scaler: Optional[dict] = None):
    """clip_grad_norm_(params, max_norm) then torch.optim.RMSprop / Adam (model/model.py:824-827, :2021-2024, :996-998).
    `state` carries the optimizer state between steps.  Returns the total gradient norm before clipping.

    scaler (dict with 'scale', 'tracker'; torch GradScaler defaults growth 2 / backoff 0.5 / interval 2000) selects the
    float16 branch (model/model.py:970-989): the loss is scaled by S before backward, clip_grad_norm_ then sees S*g,
    scaler.step() unscales and skips the optimizer when a gradient is non-finite, scaler.update() moves S.  The
    gradients given here are exact (unscaled); the reference's fp16 overflow is restated on the parameter gradients,
    all of which pass through fp16 under autocast: overflow <=> S * max|g| >= 65520 (or a non-finite gradient).
    Returns S * ||g|| in that case, like clip_grad_norm_ there; scaler['skipped'] says whether the step was dropped."""
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    if scaler is not None:
        S = float(scaler["scale"])
        gmax = max(float(np.abs(g).max()) for g in grads.values())
        bad = (not np.isfinite(total)) or (not np.isfinite(gmax)) or S * gmax >= FP16_OVERFLOW
        scaler["skipped"] = bool(bad)
        if bad:
            scaler["scale"], scaler["tracker"] = S * scaler.get("backoff", 0.5), 0
            return S * total
        scaler["tracker"] = scaler.get("tracker", 0) + 1
        if scaler["tracker"] >= scaler.get("interval", 2000):
            scaler["scale"], scaler["tracker"] = S * scaler.get("growth", 2.0), 0
        total *= S
        grads = {k: g.astype(np.float64) for k, g in grads.items()}
        coef = min(1.0, max_norm / (total + 1e-6)) if max_norm and max_norm > 0 else 1.0
    else:
        coef = min(1.0, max_norm / (total + 1e-6)) if max_norm and max_norm > 0 else 1.0
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    for k, g in grads.items():
        g = g.astype(np.float64).reshape(sd[k].shape) * coef
        if optimizer == "rmsprop":
            e = 1e-8 if eps is None else eps
            sq = state.get(("sq", k), 0.0) * alpha + (1 - alpha) * g * g
            state[("sq", k)] = sq
            sd[k] = (sd[k] - lr * g / (np.sqrt(sq) + e)).astype(np.float32)
        else:
            e = 1e-8 if eps is None else eps
            m = state.get(("m", k), 0.0) * betas[0] + (1 - betas[0]) * g
            v = state.get(("v", k), 0.0) * betas[1] + (1 - betas[1]) * g * g
            state[("m", k)], state[("v", k)] = m, v
            denom = np.sqrt(v) / np.sqrt(1 - betas[1] ** t) + e
            sd[k] = (sd[k] - lr / (1 - betas[0] ** t) * m / denom).astype(np.float32)
    return total


def laff_train_step(sd: dict, vis_in: Mapping[str, np.ndarray], txt_in: Mapping[str, np.ndarray], state: dict, heads: int,
                    vis_no_transform: Sequence[str], optimizer="rmsprop", lr=1e-4, grad_clip=2.0, margin=0.2, adam_eps=1e-4,
                    with_ave=False, mul=False, loss_kind="mrl", gru_tokens: Optional[Sequence[np.ndarray]] = None):
    """W2VVPP_MultiHeadAttention.forward(train_data) (model/model.py:964-1001 with :2021-2048): one step on the full
    model state dict (keys 'vis_net.*' / 'txt_net.*'), dropout 0.  Returns (loss, clipped gradients, grad norm)."""
    vfe = [(n, x) for n, x in vis_in.items()]
    vpre = {n: "vis_net.VisMutiTransformNet.%s." % n for n in vis_in}
    gru_prefix = "txt_net.encoder.rnn_encoder."
    if gru_tokens is not None:  # the GRU front-end trains with the model (token ids instead of a precomputed 'gru' feature)
        txt_in = dict(txt_in, gru=gru_train(gru_tokens, sd, gru_prefix))
    tfe = [(TXT_FEATURE_KEY[e], txt_in[TXT_FEATURE_KEY[e]]) for e in TXT_ENCODER_ORDER if TXT_FEATURE_KEY[e] in txt_in]
    tpre = {TXT_FEATURE_KEY[e]: "txt_net.transform_layer.%s_transform." % e for e in TXT_ENCODER_ORDER}
    t_emb, tc = fusion_train_forward(tfe, sd, tpre, "txt_net.attention_layer.", heads, ["clip"], with_ave=with_ave, mul=mul)
    v_emb, vc = fusion_train_forward(vfe, sd, vpre, "vis_net.attention_layer.", heads, vis_no_transform, with_ave=with_ave, mul=mul)
    if loss_kind == "dsl":                                                        # model/model.py:1995-1996, :2036-2038
        loss, d_txt, d_vis = 0.0, np.zeros_like(t_emb), np.zeros_like(v_emb)
        for h in range(heads):
            l, a, b = dual_softmax_loss(t_emb[:, h], v_emb[:, h], 1000.0, want_grad=True)
            loss += l
            d_txt[:, h], d_vis[:, h] = a, b
    else:
        loss, d_txt, d_vis = multi_head_loss(t_emb, v_emb, margin, True, "sum", "t2i", want_grad=True)
    grads = {}
    grads.update(fusion_train_backward(tc, d_txt, sd))
    grads.update(fusion_train_backward(vc, d_vis, sd))
    if gru_tokens is not None:
        grads.update(gru_train(gru_tokens, sd, gru_prefix, tc["dx"]["gru"]))
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    coef = min(1.0, grad_clip / (total + 1e-6)) if grad_clip and grad_clip > 0 else 1.0
    clipped = {k: (g * coef) for k, g in grads.items()}
    clip_and_step(sd, grads, state, optimizer, lr, grad_clip, eps=(adam_eps if optimizer == "adam" else None))
    return float(loss), clipped, total


def gru_train(idx_vecs: Sequence[np.ndarray], sd: Mapping[str, np.ndarray], prefix: str, dout: Optional[np.ndarray] = None):
    """GruTxtEncoder in train mode with 'mean' pooling (model/model.py:340-369) in float64.  Without dout: the pooled
    features [B, H].  With dout [B, H]: the gradients of we.weight / rnn.{weight,bias}_{ih,hh}_l0 by backward through time."""
    we, w_ih, w_hh = (sd[prefix + k].astype(np.float64) for k in ("we.weight", "rnn.weight_ih_l0", "rnn.weight_hh_l0"))
    b_ih, b_hh = sd[prefix + "rnn.bias_ih_l0"].astype(np.float64), sd[prefix + "rnn.bias_hh_l0"].astype(np.float64)
    H = w_hh.shape[1]
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    outs, tapes = [], []
    for ids in idx_vecs:
        h = np.zeros(H)
        tape = []
        for tok in ids:
            x = we[tok]
            gi, gh = w_ih @ x + b_ih, w_hh @ h + b_hh
            r, z = sig(gi[:H] + gh[:H]), sig(gi[H:2 * H] + gh[H:2 * H])
            n = np.tanh(gi[2 * H:] + r * gh[2 * H:])
            hn = (1 - z) * n + z * h
            tape.append((tok, x, h, r, z, n, gh[2 * H:]))
            h = hn
        tapes.append(tape)
        hs = [(1 - t[4]) * t[5] + t[4] * t[2] for t in tape]
        outs.append(np.mean(hs, axis=0))
    if dout is None:
        return np.stack(outs)
    g = {k: np.zeros_like(v) for k, v in (("we.weight", we), ("rnn.weight_ih_l0", w_ih), ("rnn.weight_hh_l0", w_hh),
                                          ("rnn.bias_ih_l0", b_ih), ("rnn.bias_hh_l0", b_hh))}
    for b, tape in enumerate(tapes):
        L = len(tape)
        dh_next = np.zeros(H)
        for tok, x, hp, r, z, n, ghn in reversed(tape):
            dh = dh_next + dout[b] / L
            dn, dz = dh * (1 - z), dh * (hp - n)
            dpn, dpz = dn * (1 - n * n), dz * z * (1 - z)
            dpr = dpn * ghn * r * (1 - r)
            dgi = np.concatenate([dpr, dpz, dpn])
            dgh = np.concatenate([dpr, dpz, dpn * r])
            g["rnn.weight_ih_l0"] += np.outer(dgi, x)
            g["rnn.weight_hh_l0"] += np.outer(dgh, hp)
            g["rnn.bias_ih_l0"] += dgi
            g["rnn.bias_hh_l0"] += dgh
            g["we.weight"][tok] += w_ih.T @ dgi
            dh_next = dh * z + w_hh.T @ dgh
    return {prefix + k: v for k, v in g.items()}


def laff_ml_train_step(sd: dict, vis_in: Mapping[str, np.ndarray], frames: np.ndarray, frame_feat: str,
                       txt_in: Mapping[str, np.ndarray], state: dict, heads: int, optimizer="rmsprop", lr=1e-4, grad_clip=2.0,
                       margin=0.2, adam_eps=1e-8, scaler: Optional[dict] = None):
    """W2VVPP_MutiVisFrameFeat ('FrameLAFF' / LAFF-ml) training step: as laff_train_step with the frame-level block in
    front of the video net (model/model.py:2147-2190) and BatchNorm on every projected feature."""
    wkey = "vis_net.frame_attention.%s.0.embedding_common.0." % frame_feat
    pooled = frame_attention_train(frames, sd[wkey + "weight"].reshape(-1))
    vfe = [(n, x) for n, x in vis_in.items()] + [(frame_feat, pooled)]
    vpre = {n: "vis_net.%s." % n for n, _ in vfe}
    tfe = [(TXT_FEATURE_KEY[e], txt_in[TXT_FEATURE_KEY[e]]) for e in TXT_ENCODER_ORDER if TXT_FEATURE_KEY[e] in txt_in]
    tpre = {TXT_FEATURE_KEY[e]: "txt_net.transform_layer.%s_transform." % e for e in TXT_ENCODER_ORDER}
    t_emb, tc = fusion_train_forward(tfe, sd, tpre, "txt_net.attention_layer.", heads, ["clip"])
    v_emb, vc = fusion_train_forward(vfe, sd, vpre, "vis_net.vis_attention_layer.", heads, [frame_feat])
    loss, d_txt, d_vis = multi_head_loss(t_emb, v_emb, margin, True, "sum", "t2i", want_grad=True)
    grads = {}
    grads.update(fusion_train_backward(tc, d_txt, sd))
    grads.update(fusion_train_backward(vc, d_vis, sd))
    dw, dc = frame_attention_train(frames, sd[wkey + "weight"].reshape(-1), vc["dx"][frame_feat])
    grads[wkey + "weight"] = dw.reshape(1, -1)
    grads[wkey + "bias"] = np.array([dc])
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    if scaler is not None:
        total *= float(scaler["scale"])
    coef = min(1.0, grad_clip / (total + 1e-6)) if grad_clip and grad_clip > 0 else 1.0
    clipped = {k: (g * coef) for k, g in grads.items()}
    clip_and_step(sd, grads, state, optimizer, lr, grad_clip, eps=(adam_eps if optimizer == "adam" else None), scaler=scaler)
    return float(loss), clipped, total


# ----------------------------------------------------------------------------------------------------------------
# text front-end
# ----------------------------------------------------------------------------------------------------------------
def tokenize(input_str: str, clean: bool = True, remove_stopword: bool = False, stopwords=()) -> list:
    """textlib.py:27-47, English branch."""
    import re
    sent = input_str
    if clean:
        sent = sent.replace("\r", " ")
        sent = re.sub(r"[^A-Za-z0-9]", " ", sent).strip().lower()
    tokens = sent.split()
    if remove_stopword:
        tokens = [x for x in tokens if x not in stopwords]
    return tokens


def bow_encoding(words: Sequence[str], word2idx: Mapping[str, int]) -> np.ndarray:
    """BowVec._encoding (txt2vec.py:56-63)."""
    vec = np.zeros(len(word2idx))
    for w in words:
        idx = word2idx.get(w, -1)
        if idx >= 0:
            vec[idx] += 1
    return vec


def w2v_encoding(words: Sequence[str], name2index: Mapping[str, int], table: np.ndarray) -> np.ndarray:
    """W2Vec._encoding (txt2vec.py:97-104) over BigFile.read (bigfile.py:197-220): distinct known words, in file
    order, as float64 lists; mean over them, zeros if none."""
    idx = sorted({name2index[w] for w in words if w in name2index})
    if not idx:
        return np.zeros(table.shape[1])
    return np.array([table[i].tolist() for i in idx]).mean(axis=0)


def index_encoding(words: Sequence[str], word2idx: Mapping[str, int]) -> np.ndarray:
    """IndexVec (txt2vec.py:121-128) with the 'gru' vocabulary's <unk> rule (textlib.py:102-109)."""
    words = ["<start>"] + list(words) + ["<end>"]
    return np.array([word2idx[w] if w in word2idx else word2idx["<unk>"] for w in words])


def gru_encoder(idx_vecs: Sequence[np.ndarray], we: np.ndarray, w_ih: np.ndarray, w_hh: np.ndarray, b_ih: np.ndarray,
                b_hh: np.ndarray, pooling: str = "mean") -> np.ndarray:
    """GruTxtEncoder.forward (model/model.py:340-387): embedding, one-layer nn.GRU (gate order r, z, n; h0 = 0) run over
    each sequence's own length (pack_padded_sequence), then mean / last / mean_last pooling."""
    H = w_hh.shape[1]
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    outs = []
    for ids in idx_vecs:
        h = np.zeros(H, dtype=np.float32)
        hs = []
        for tok in ids:
            x = we[tok].astype(np.float32)
            gi = w_ih @ x + b_ih
            gh = w_hh @ h + b_hh
            r = sig(gi[:H] + gh[:H])
            z = sig(gi[H:2 * H] + gh[H:2 * H])
            n = np.tanh(gi[2 * H:] + r * gh[2 * H:])
            h = ((1.0 - z) * n + z * h).astype(np.float32)
            hs.append(h)
        hs = np.stack(hs)
        mean, last = hs.mean(axis=0), hs[-1]
        outs.append(mean if pooling == "mean" else last if pooling == "last" else np.concatenate([mean, last]))
    return np.stack(outs).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------
# predictor / validate: id-based ground truth, both directions, result writers  
# ----------------------------------------------------------------------------------------------------------------
def sorted_desc(scores: np.ndarray) -> np.ndarray:
    """`inds[index][::-1]` of predictor.py:232-239 under the documented tie rule (stable argsort, reversed)."""
    return np.argsort(scores, axis=1, kind="stable")[:, ::-1]


def predictor_t2v_eval(t2i: np.ndarray, txt_ids: Sequence[str], vis_ids: Sequence[str]):
    """predictor.py:236-246 / trainer.py:582-599 — label_matrix[i, p] = 1 where the p-th ranked video id equals
    txt_id.split('#')[0]; metrics = evaluation.eval(label_matrix).  Returns (metrics7, label_matrix)."""
    order = sorted_desc(t2i)
    vis = np.array(vis_ids)
    label = np.zeros(order.shape)
    for i in range(order.shape[0]):
        label[i][np.where(vis[order[i]] == txt_ids[i].split("#")[0])[0]] = 1
    return eval_label_matrix(label), label


def predictor_v2t_eval(t2i: np.ndarray, txt_ids: Sequence[str], vis_ids: Sequence[str]):
    """predictor.py:262-270 — rows = videos of t2i.T, ground truths = every caption whose id prefix is the video id."""
    i2t = t2i.T
    order = sorted_desc(i2t)
    cap_vid = np.array([t.split("#")[0] for t in txt_ids])
    label = np.zeros(order.shape)
    for v in range(order.shape[0]):
        label[v][np.where(cap_vid[order[v]] == vis_ids[v])[0]] = 1
    return eval_label_matrix(label), label


def writer_topk(n_vis: int, threshold: int) -> int:
    """predictor.py:55-58, :64 — `ind = inds[index][::-1][0:TopK]` with TopK = Threshold if len(vis_ids) >= Threshold
    else -1: below the threshold the slice 0:-1 drops the last (lowest-ranked) video."""
    return threshold if n_vis >= threshold else max(n_vis - 1, 0)


def txt2video_lines(t2i: np.ndarray, txt_ids: Sequence[str], vis_ids: Sequence[str], threshold: int = 2000):
    """The lines txt2video_write_to_file writes (predictor.py:62-67): '<txt_id> <vis_id> <score> <vis_id> <score> ...'
    with scores printed by '%s' of the float32 matrix entry."""
    order = sorted_desc(t2i)
    k = writer_topk(len(vis_ids), threshold)
    lines = []
    for i in range(order.shape[0]):
        ind = order[i][:k]
        lines.append(txt_ids[i] + " " + " ".join([vis_ids[j] + " %s" % t2i[i][j] for j in ind]))
    return lines


def t2v_shot_dict(t2i: np.ndarray, txt_ids: Sequence[str], vis_ids: Sequence[str], captions: Mapping[str, str],
                  threshold: int = 500):
    """The dict pickled to t2v.pkl (predictor.py:68-86)."""
    order = sorted_desc(t2i)
    k = writer_topk(len(vis_ids), threshold)
    out = {}
    for i in range(order.shape[0]):
        ind = order[i][:k]
        out[txt_ids[i]] = {"query": captions[txt_ids[i]], "rank_list": [vis_ids[j] for j in ind],
                           "sim_value": [t2i[i][j] for j in ind]}
    return out


# ----------------------------------------------------------------------------------------------------------------
# loss
# ----------------------------------------------------------------------------------------------------------------
def _hinge(scores: np.ndarray, margin, max_violation: bool, cost_style: str, direction: str, want_grad: bool):
    """Shared by MarginRankingLoss (loss.py:99-135) and MarginRankingLossWithScore (loss.py:164-200).
    Returns (loss, dLoss/dscores)."""
    dt = scores.dtype.type
    B = scores.shape[0]
    diag = np.diag(scores).reshape(B, 1)
    eye = np.eye(B, dtype=bool)
    loss = dt(0)
    g = np.zeros_like(scores)
    parts = []
    if direction in ("i2t", "bidir"):
        parts.append(("s", np.where(eye, 0, np.maximum(dt(margin) + scores - diag, 0))))        # d1: diag[i] per row
    if direction in ("t2i", "bidir"):
        parts.append(("im", np.where(eye, 0, np.maximum(dt(margin) + scores - diag.T, 0))))     # d2: diag[j] per col
    for kind, cost in parts:
        if max_violation:
            red = cost.max(1) if kind == "s" else cost.max(0)                                      # loss.py:121-125
            arg = cost.argmax(1) if kind == "s" else cost.argmax(0)
            denom = dt(B) if cost_style != "sum" else dt(1)
            loss = loss + red.sum(dtype=scores.dtype) / denom
            if want_grad:
                for t in range(B):
                    if red[t] > 0:
                        if kind == "s":
                            g[t, arg[t]] += 1 / denom
                            g[t, t] -= 1 / denom
                        else:
                            g[arg[t], t] += 1 / denom
                            g[t, t] -= 1 / denom
        else:
            denom = dt(B * B) if cost_style != "sum" else dt(1)
            loss = loss + cost.sum(dtype=scores.dtype) / denom
            if want_grad:
                viol = (cost > 0).astype(scores.dtype) / denom
                g += viol
                if kind == "s":
                    g[np.arange(B), np.arange(B)] -= viol.sum(1)
                else:
                    g[np.arange(B), np.arange(B)] -= viol.sum(0)
    return loss, g


def _l2norm_backward(x: np.ndarray, ghat: np.ndarray, eps) -> np.ndarray:
    """d/dx of x / (||x|| + eps) contracted with ghat (autograd of loss.py:11-12)."""
    n = np.sqrt((x * x).sum(1, keepdims=True))
    den = n + eps
    xhat = x / den
    dot = (xhat * ghat).sum(1, keepdims=True)
    return ghat / den - xhat * dot / np.where(n > 0, n, 1)


def margin_ranking_loss(s: np.ndarray, im: np.ndarray, margin=0.0, max_violation=False, cost_style="sum",
                        direction="bidir", want_grad: bool = False):
    """MarginRankingLoss.forward(s, im) (loss.py:95-135): scores = cosine_sim(im, s) so rows = images/videos and
    columns = sentences.  Returns loss or (loss, d_s, d_im)."""
    dt = s.dtype.type
    eps = dt(1e-13) + dt(1e-14)
    s_hat, im_hat = l2norm(s), l2norm(im)
    scores = im_hat @ s_hat.T
    loss, g = _hinge(scores, margin, max_violation, cost_style, direction, want_grad)
    if not want_grad:
        return loss
    d_im_hat = g @ s_hat
    d_s_hat = g.T @ im_hat
    return loss, _l2norm_backward(s, d_s_hat, eps), _l2norm_backward(im, d_im_hat, eps)


def multi_head_loss(txt_embs: np.ndarray, vis_embs: np.ndarray, margin=0.2, max_violation=True, cost_style="sum",
                    direction="t2i", want_grad: bool = False):
    """W2VVPP.compute_loss 3-D branch / W2VVPP_MultiHeadAttention.compute_loss multi_space branch
    (model/model.py:852-862, :2036-2038): sum over heads of criterion(txt[:, h, :], vis[:, h, :])."""
    total = txt_embs.dtype.type(0)
    d_txt = np.zeros_like(txt_embs)
    d_vis = np.zeros_like(vis_embs)
    for h in range(vis_embs.shape[1]):
        r = margin_ranking_loss(np.ascontiguousarray(txt_embs[:, h, :]), np.ascontiguousarray(vis_embs[:, h, :]), margin,
                                max_violation, cost_style, direction, want_grad)
        if want_grad:
            total = total + r[0]
            d_txt[:, h, :], d_vis[:, h, :] = r[1], r[2]
        else:
            total = total + r
    return (total, d_txt, d_vis) if want_grad else total


def dual_softmax_loss(s: np.ndarray, im: np.ndarray, temp=1000.0, want_grad: bool = False):
    """DualSoftmaxLoss.forward(s, im, temp) (loss.py:291-310) on [B, d] embeddings, in float64; with want_grad also
    (d loss / d s, d loss / d im) through cosine_sim's normalisation."""
    s64, im64 = s.astype(np.float64), im.astype(np.float64)
    eps = 1e-13 + 1e-14
    ns, ni = np.sqrt((s64 ** 2).sum(1, keepdims=True)), np.sqrt((im64 ** 2).sum(1, keepdims=True))
    sh, ih = s64 / (ns + eps), im64 / (ni + eps)
    sim = sh @ ih.T
    B = sim.shape[0]

    def term(A):
        P0 = softmax(A / temp, axis=0)
        M = A * P0 * B
        L = M - M.max(1, keepdims=True)
        logp = L - np.log(np.exp(L).sum(1, keepdims=True))
        loss = -np.trace(logp)
        dM = np.exp(logp) - np.eye(B)
        dP0 = B * A * dM
        dA = B * P0 * dM + P0 * (dP0 - (P0 * dP0).sum(0, keepdims=True)) / temp
        return loss, dA

    l1, d1 = term(sim)
    l2, d2 = term(sim.T)
    loss = (l1 + l2) / 2
    if not want_grad:
        return loss
    dsim = (d1 + d2.T) / 2
    gs, gi = dsim @ ih, dsim.T @ sh
    return loss, _l2norm_backward(s64, gs, eps), _l2norm_backward(im64, gi, eps)


def margin_ranking_loss_with_score(score: np.ndarray, margin=0.0, max_violation=False, cost_style="sum",
                                   direction="bidir", want_grad: bool = False):
    """MarginRankingLossWithScore.forward(score) (loss.py:161-200)."""
    loss, g = _hinge(score, margin, max_violation, cost_style, direction, want_grad)
    return (loss, g) if want_grad else loss


# ----------------------------------------------------------------------------------------------------------------
# whole-path driver used as the CPU baseline (bench.py) — the reference's predictor path on embeddings
# ----------------------------------------------------------------------------------------------------------------
def retrieve_cpu(q_emb: np.ndarray, g_emb: np.ndarray, gt: np.ndarray, heads: int, k: int = 10, chunk: int = 100000,
                 threads: int = 1):
    """Reference evaluation path for Q queries against a V-video gallery of fused embeddings, restated for bounded
    memory: get_txt2vis_matrix (model/model.py:1003-1016) per gallery chunk, np.argsort per row (predictor.py:232),
    ground-truth position (predictor.py:239-244), metrics (evaluation.py:81-89).  The gallery is processed in
    `chunk`-video pieces (the reference's predict() tiles by loader batch, model/model.py:1064-1073) and rows are
    sorted on `threads` host threads.  Returns (rank0 [Q], topk_idx [Q, k], metrics)."""
    Q = q_emb.shape[0]
    V = g_emb.shape[0]
    H = heads
    q3 = q_emb.reshape(Q, H, -1)
    scores = np.empty((Q, V), dtype=np.float32)
    for s in range(0, V, chunk):
        e = min(V, s + chunk)
        scores[:, s:e] = txt2vis_matrix(q3, g_emb[s:e].reshape(e - s, H, -1))
    rank0 = np.empty(Q, dtype=np.int64)
    topk = np.empty((Q, k), dtype=np.int64)

    def work(lo, hi):
        inds = np.argsort(scores[lo:hi], axis=1)
        for i in range(lo, hi):
            ind = inds[i - lo][::-1]
            rank0[i] = np.where(ind == gt[i])[0][0]
            topk[i] = ind[:k]

    step = max(1, (Q + threads - 1) // threads)
    if threads <= 1:
        work(0, Q)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda lo: work(lo, min(Q, lo + step)), range(0, Q, step)))
    return rank0, topk, metrics_from_rank0(rank0)
