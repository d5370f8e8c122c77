"""Seeded synthetic inputs and parameters for the LAFF hot path (shared by tests, golden generation and bench.py).

Everything is generated with ``numpy.random.RandomState`` keyed by (seed, name) so that any consumer can regenerate a
tensor by name without depending on generation order (the reference's constructors consume torch RNG while building
all 16 attention variants, model/model.py:95-206, so seeds alone cannot reproduce its own initialisation).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Mapping, Tuple

import numpy as np

# Feature dimensions: pinned by the reference unless marked assumed.
DIMS = {
    "clip": 512,     # configs/laff.py:36 clip_opt['size']
    "gru": 1024,     # configs/base_config.py:38 rnn_size
    "w2v": 500,      # configs/base_config.py:37, trainer.py:190
    "bow": 3981,     # data/vocab_tgif.zip bow_nsw_5 vocabulary
    "x3d": 2048,     # assumed (public backbone)
    "ircsn": 2048,   # assumed
    "tf": 768,       # assumed (TimeSformer)
    "c3d": 2048,     # assumed
}

# reference feature names (configs/laff.py:54-65, FrameLaff...:63-66,104-112)
VIS_CLIP_FT = "clip_finetune_8frame_uniform_1103"
VIS_TF = "HowTo100M_TimeSformer_divST_96x4_224"
VIS_X3D = "X3D_L"
VIS_IRCSN = "mean_irCSN_152_ig65m_from_scratch"
VIS_C3D = "mean_C3d_resneXt101_16f"
VIS_FRAME = "Frame_clip_finetune_8frame_uniform_1103"


def rng_for(seed: int, name: str) -> np.random.RandomState:
    return np.random.RandomState((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 32))


def feature(seed: int, name: str, rows: int, dim: int, kind: str = "dense") -> np.ndarray:
    """One synthetic feature matrix [rows, dim] float32.

    dense: N(0,1); relu: max(N(0,1),0) (pooled CNN features); bow: counts of 8 random vocabulary ids per row.
    """
    r = rng_for(seed, "feat/" + name)
    if kind == "bow":
        out = np.zeros((rows, dim), dtype=np.float32)
        ids = r.randint(0, dim, size=(rows, 8))
        for i in range(rows):
            np.add.at(out[i], ids[i], 1.0)
        return out
    x = r.standard_normal((rows, dim)).astype(np.float32)
    if kind == "relu":
        x = np.maximum(x, 0.0)
    return x


def param(seed: int, key: str, shape: Tuple[int, ...], omega: float = 1.0) -> np.ndarray:
    """Synthetic value of one state_dict entry, chosen by the reference parameter name."""
    r = rng_for(seed, "param/" + key)
    shape = tuple(int(s) for s in shape)
    if key.endswith("num_batches_tracked"):
        return np.asarray(1, dtype=np.int64)
    if key.endswith("global_emb_weight_net.weight"):
        return np.full(shape, omega, dtype=np.float32)
    if ".bn1." in key or key.endswith(("bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var")):
        if key.endswith("running_var"):
            return r.uniform(0.5, 2.0, shape).astype(np.float32)
        if key.endswith("running_mean"):
            return (0.5 * r.standard_normal(shape)).astype(np.float32)
        if key.endswith("weight"):
            return r.uniform(0.5, 1.5, shape).astype(np.float32)
        return (0.1 * r.standard_normal(shape)).astype(np.float32)
    if "layer_norm" in key:
        return (np.ones(shape) if key.endswith("weight") else np.zeros(shape)).astype(np.float32)
    if "embedding_common" in key:
        dh = shape[-1] if key.endswith("weight") else 1
        if key.endswith("weight"):
            return r.uniform(-3.0, 3.0, shape).astype(np.float32) / np.float32(np.sqrt(dh))
        return r.uniform(-0.1, 0.1, shape).astype(np.float32)
    if key.endswith("fc1.weight") or (len(shape) == 2 and key.endswith("weight")):
        bound = np.sqrt(6.0 / (shape[0] + shape[1]))  # xavier_uniform_ (model/model.py:55)
        return r.uniform(-bound, bound, shape).astype(np.float32)
    if key.endswith("bias"):
        return (0.05 * r.standard_normal(shape)).astype(np.float32)
    return (0.1 * r.standard_normal(shape)).astype(np.float32)


def state_dict(seed: int, shapes: Mapping[str, Tuple[int, ...]], omega: float = 1.0) -> Dict[str, np.ndarray]:
    return {k: param(seed, k, s, omega) for k, s in shapes.items()}


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float32 (numpy has no bf16)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


def unit_heads(x: np.ndarray, heads: int) -> np.ndarray:
    rows = x.shape[0]
    y = x.reshape(rows, heads, -1).astype(np.float32)
    y = y / np.sqrt((y * y).sum(2, keepdims=True))
    return y.reshape(rows, -1)


def sigma_for_recall(V: int, D: int, z_shift: float = -0.52) -> float:
    """Noise level for which R@1 is ~30% against V unit-norm random distractors: the positive's mean score
    1/sqrt(1+sigma^2) sits z_shift standard deviations from the expected maximum negative score
    (Gumbel approximation of the maximum of V standard normals, scaled by 1/sqrt(D))."""
    lv = np.log(float(max(V, 3)))
    z = np.sqrt(2 * lv) - (np.log(lv) + np.log(4 * np.pi)) / (2 * np.sqrt(2 * lv))
    t = max(1e-3, (z + z_shift) / np.sqrt(D))
    return float(np.sqrt(max(1.0 / (t * t) - 1.0, 0.0)))


def retrieval_embeddings(seed: int, Q: int, V: int, heads: int = 8, head_dim: int = 512, sigma: float = 1.2):
    """C5-style synthetic embeddings: gallery = unit-norm noise per head, gt(i) = (i*97) mod V,
    query = normalize(gallery[gt] + sigma * unit noise).  Returns (q [Q,D], g [V,D], gt [Q]) float32 / int64."""
    D = heads * head_dim
    g = unit_heads(rng_for(seed, "emb/gallery").standard_normal((V, D)).astype(np.float32), heads)
    gt = (np.arange(Q, dtype=np.int64) * 97) % V
    noise = unit_heads(rng_for(seed, "emb/query").standard_normal((Q, D)).astype(np.float32), heads)
    q = unit_heads(g[gt] + sigma * noise, heads)
    return q, g, gt


# Dimensions of the trained-checkpoint fixture (tests/golden/make_golden_trained.py): the shipped head size d_h = 512 (so
# the single-kernel fusion path and the 16-bit similarity sweep are the ones exercised) with H = 4 heads and narrow input
# features, which keeps the reference-trained checkpoint at ~4 MB.
TRAINED_DIMS = {"clip": 512, "gru": 96, "bow": 128, "w2v": 60, "x3d": 96, "ircsn": 96, "tf": 64, "c3d": 96}
TRAINED_HEADS, TRAINED_D, TRAINED_LATENT = 4, 2048, 64


def latent_collection(seed: int, n: int, dims: Mapping[str, int] = None, map_seed: int = 4242, latent: int = TRAINED_LATENT,
                      vis_noise: float = 0.6, cap_noise: float = 0.5, txt_noise: float = 0.6):
    """n (video, caption) pairs that share a latent factor (SURVEY §8d "trained fixture"): z_i ~ N(0, I_64); every video
    feature is a fixed random linear map of z_i plus noise (pooled-CNN-like features pass through a ReLU); the caption's
    latent is z_i plus noise and every text feature a fixed random map of it plus noise; the BoW vector marks the 8
    vocabulary entries with the largest (map + Gumbel noise) score.  The maps depend on `map_seed` only, so collections
    of different `seed` (train / test) are draws from the same distribution.  Ground truth: caption i <-> video i.
    Returns (vis_feats {reference feature name: [n, d]}, txt_feats {'gru','bow','w2v','clip': [n, d]})."""
    d = dict(TRAINED_DIMS if dims is None else dims)
    z = rng_for(seed, "latent/z").standard_normal((n, latent)).astype(np.float32)
    zc = z + np.float32(cap_noise) * rng_for(seed, "latent/cap").standard_normal((n, latent)).astype(np.float32)

    def mapped(zz, name, dim, noise, relu):
        m = rng_for(map_seed, "map/" + name).standard_normal((latent, dim)).astype(np.float32) / np.float32(np.sqrt(latent))
        x = zz @ m + np.float32(noise) * rng_for(seed, "noise/" + name).standard_normal((n, dim)).astype(np.float32)
        return np.maximum(x, 0.0).astype(np.float32) if relu else x.astype(np.float32)

    vis = {VIS_CLIP_FT: mapped(z, "v/clip", d["clip"], vis_noise, False), VIS_TF: mapped(z, "v/tf", d["tf"], vis_noise, True),
           VIS_X3D: mapped(z, "v/x3d", d["x3d"], vis_noise, True), VIS_IRCSN: mapped(z, "v/ircsn", d["ircsn"], vis_noise, True)}
    txt = {"gru": mapped(zc, "t/gru", d["gru"], txt_noise, False), "w2v": mapped(zc, "t/w2v", d["w2v"], txt_noise, False),
           "clip": mapped(zc, "t/clip", d["clip"], txt_noise, False)}
    score = mapped(zc, "t/bow", d["bow"], 0.0, False) + 0.5 * rng_for(seed, "noise/bow").gumbel(size=(n, d["bow"])).astype(np.float32)
    ids = np.argsort(-score, axis=1, kind="stable")[:, :8]
    bow = np.zeros((n, d["bow"]), dtype=np.float32)
    np.put_along_axis(bow, ids, 1.0, axis=1)
    txt["bow"] = bow
    return vis, txt
