"""End-to-end parity at the BASELINE.json configurations that are parity cases rather than bench lines:

  C1  LAFF eval, MSR-VTT-1k-shaped: 1000 queries x 1000 videos (clip-ft + x3d + ircsn video, clip + bow + w2v + gru text)
  C2  MV-test3k-shaped: 2990 x 2990 with the shipped 4-feature video set, fused encode + sim + rank on one GPU
  C4  frame-level LAFF (LAFF-ml video side) over 32-frame CLIP features, TGIF-test-shaped gallery (11 360 videos)

Raw synthetic features go in, R@K / MedR come out; the oracle runs the reference's whole pipeline on the CPU in fp32
(fusion nets -> get_txt2vis_matrix -> argsort rank -> metrics).  Inputs and FC weights are bf16-representable, so the
only differences are accumulation order and the bf16 rounding of the embeddings fed to the similarity GEMM:
  * fused embeddings: T1, <= 2e-6;
  * 'bf16x3' similarity (near-fp32 products): ranks identical to the oracle except for queries with a competitor within
    the numerical noise of s_gt, which are enumerated; R@1/5/10 within one query, MedR identical;
  * 'bf16' similarity (the default): T2, reported — score error <= 2e-3, R@K within 1 point.
"""
import numpy as np
import pytest
import torch

from helpers import load_numpy_state, max_abs
from laff_b200 import config as cfg
from laff_b200 import loss as L
from laff_b200 import model as M
from laff_b200 import ops, synth
from laff_b200.retrieval import GalleryIndex
from oracle import laff_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore():
    L.set_precision("bf16")
    yield
    L.set_precision("bf16")


LATENT = 64
ALIAS = {synth.VIS_CLIP_FT: "clip"}  # CLIP video and text features live in one space: same latent map


def latent_map(seed, name, d):
    return synth.rng_for(seed, "map/" + ALIAS.get(name, name)).standard_normal((LATENT, d)).astype(np.float32) / np.sqrt(LATENT)


def latent_features(seed, n, dims, kinds, noise=5.0):
    """Every feature of item i is a fixed random map of a shared latent z_i plus noise; bf16-representable."""
    r = synth.rng_for(seed, "latent")
    z = r.standard_normal((n, LATENT)).astype(np.float32)
    out = {}
    for name, d in dims.items():
        A = latent_map(seed, name, d)
        x = z @ A + noise * synth.rng_for(seed + 1, "noise/" + name).standard_normal((n, d)).astype(np.float32)
        if kinds.get(name) == "relu":
            x = np.maximum(x, 0)
        if kinds.get(name) == "bow":
            x = np.round(np.maximum(x, 0))
        out[name] = synth.bf16_round(x) if kinds.get(name) != "raw" else x.astype(np.float32)
    return out


TXT_KEY = {"rnn_encoder": "gru", "bow_encoder": "bow", "w2v_encoder": "w2v"}


def make_nets(c, seed):
    """Synthetic parameters standing in for a trained checkpoint (random weights give chance-level retrieval): every FC
    maps its feature back to the shared latent and on through one common projection P, W_l = P pinv(A_l)^T, so that
    text and video embeddings of the same item correlate.  FC weights are bf16-representable."""
    vis = M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0])
    txt = M.MultiScaleTxtEncoderAttention(c)
    P = synth.rng_for(seed, "P").standard_normal((4096, LATENT)).astype(np.float32) * 0.8
    sds = []
    for net, s in ((vis, seed), (txt, seed + 1)):
        sd = {k: synth.param(s, k, tuple(v.shape)) for k, v in net.state_dict().items()}
        for k in sd:
            if k.endswith("fc1.weight"):
                feat = k.split(".")[-3]
                feat = TXT_KEY.get(feat.replace("_transform", ""), feat)
                A = latent_map(seed, feat, sd[k].shape[1])
                sd[k] = synth.bf16_round((P @ np.linalg.pinv(A).T).astype(np.float32))
        load_numpy_state(net, sd)
        sds.append(sd)
    return vis.cuda().eval(), txt.cuda().eval(), sds[0], sds[1]


def run_config(n, vis_names, seed):
    dims = dict(synth.DIMS)
    c = cfg.laff_config(4096, 8, dims)
    c.vis_fc_layers = [{k: v for k, v in c.vis_fc_layers[0].items() if k in vis_names}, 4096]
    vis, txt, vsd, tsd = make_nets(c, seed)
    vdims = c.vis_fc_layers[0]
    vin = latent_features(seed, n, vdims, {k: ("raw" if k == synth.VIS_CLIP_FT else "relu") for k in vdims})
    tdims = {"gru": dims["gru"], "bow": dims["bow"], "w2v": dims["w2v"], "clip": dims["clip"]}
    tin = latent_features(seed, n, tdims, {"bow": "bow", "clip": "raw"})
    # --- oracle: the reference pipeline in fp32 on the CPU
    ov, _ = O.vis_net_forward(vin, vsd, [synth.VIS_CLIP_FT], 8)
    ot, _ = O.txt_net_forward(tin, tsd, ["CLIP_encoder"], 8)
    s_ref = O.txt2vis_matrix(ot, ov)
    gt = np.arange(n)
    r_ref = O.tie_rule_rank(s_ref, gt)
    m_ref = O.metrics_from_rank0(r_ref)
    # --- device
    v32, v16 = vis.encode({k: torch.from_numpy(x) for k, x in vin.items()}, out16_dtype=torch.bfloat16)
    t32, t16 = txt.encode({k: torch.from_numpy(x) for k, x in tin.items()}, out16_dtype=torch.bfloat16)
    assert max_abs(v32, ov) <= 2e-6 and max_abs(t32, ot) <= 2e-6
    out = {"ref": (s_ref, r_ref, m_ref)}
    gt_t = torch.arange(n, device="cuda", dtype=torch.int32)
    # default precision: bf16 embeddings through the fused sweep
    res = GalleryIndex(v16.reshape(n, -1), n, 8).search(t16.reshape(n, -1), gt_t, 10)
    out["bf16"] = (res.rank0.cpu().numpy(), res.metrics.cpu().numpy(), res.topk_idx.cpu().numpy())
    # near-fp32 similarity: 3-term split operands, same sweep kernel (K = 3 * 4096)
    q3, g3 = ops.split3_16(t32.reshape(n, -1), 0), ops.split3_16(v32.reshape(n, -1), 1)
    sgt = ops.sim_gt_scores(q3, g3, gt_t)
    cnt, tv, ti = ops.sim_rank_topk(q3, g3, sgt, gt_t, 10, scale=0.125)
    out["bf16x3"] = (cnt.cpu().numpy(), ops.rank_metrics(cnt).cpu().numpy(), ti.cpu().numpy(), tv.cpu().numpy())
    return out


def check(out, n):
    s_ref, r_ref, m_ref = out["ref"]
    assert 5 < m_ref[0] < 100                      # the synthetic task is neither trivial nor chance level
    sg = s_ref[np.arange(n), np.arange(n)][:, None]
    # bf16x3: identical except where a competitor sits within 2e-5 of s_gt (fp32 reference noise + split residual)
    r3, m3, ti3, tv3 = out["bf16x3"]
    near = (np.abs(s_ref - sg) < 2e-5).sum(1) - 1
    clean = near == 0
    assert clean.mean() > 0.3, clean.mean()
    np.testing.assert_array_equal(r3[clean], r_ref[clean])
    assert np.all(np.abs(r3[~clean] - r_ref[~clean]) <= near[~clean])
    assert abs(m3[3] - m_ref[3]) <= 1 and all(abs(m3[i] - m_ref[i]) <= 100.0 * (~clean).sum() / n + 1e-9 for i in range(3))
    top_ref = np.sort(s_ref, axis=1)[:, ::-1][:, :10]
    assert np.abs(tv3 - top_ref).max() <= 2e-5
    # bf16 (default): T2, reported
    r16, m16, _ = out["bf16"]
    moved = float((r16 != r_ref).mean())
    assert all(abs(m16[i] - m_ref[i]) <= 1.0 for i in range(3)) and abs(m16[3] - m_ref[3]) <= 1
    return moved


def test_c1_msrvtt1k_shaped():
    out = run_config(1000, [synth.VIS_CLIP_FT, synth.VIS_X3D, synth.VIS_IRCSN], seed=1234 + 1)
    moved = check(out, 1000)
    print("C1: R@1/5/10/MedR ref %s | bf16 %s | ranks moved by bf16 operands: %.1f%%" % (
        np.round(out["ref"][2][:4], 2), np.round(out["bf16"][1][:4], 2), 100 * moved))


def test_c2_mvtest3k_shaped():
    out = run_config(2990, [synth.VIS_CLIP_FT, synth.VIS_TF, synth.VIS_X3D, synth.VIS_IRCSN], seed=1234 + 2)
    moved = check(out, 2990)
    print("C2: R@1/5/10/MedR ref %s | bf16 %s | ranks moved by bf16 operands: %.1f%%" % (
        np.round(out["ref"][2][:4], 2), np.round(out["bf16"][1][:4], 2), 100 * moved))


def test_c4_frame_level_laff_tgif_shaped():
    """LAFF-ml video side: 11 360 videos x 32 frames x 512 (ragged, zero padded) + C3D / TimeSformer / X3D / irCSN,
    batch_norm=True, through VisMutiTransformNetPlusFrameFeat; embeddings vs the oracle, then a retrieval sanity check."""
    V, F = 11360, 32
    c = cfg.frame_laff_config(4096, 8, synth.DIMS)
    net = M.VisMutiTransformNetPlusFrameFeat(c)
    sd = {k: synth.bf16_round(synth.param(44, k, tuple(v.shape))) if k.endswith("fc1.weight") else synth.param(44, k, tuple(v.shape))
          for k, v in net.state_dict().items()}
    load_numpy_state(net, sd)
    net = net.cuda().eval()
    vdims = {k: v for k, v in c.vis_fc_layers[0].items() if k != synth.VIS_FRAME}
    vin = latent_features(1234 + 4, V, vdims, {k: "relu" for k in vdims})
    r = synth.rng_for(1234 + 4, "frames")
    frames = r.standard_normal((V, F, 512)).astype(np.float32)
    lens = r.randint(8, F + 1, size=V)
    lens[0] = F
    frames[np.arange(F)[None, :] >= lens[:, None]] = 0
    emb32, emb16 = net.encode({k: torch.from_numpy(x) for k, x in vin.items()},
                              {"mask_tensor": torch.from_numpy((np.arange(F)[None, :] < lens[:, None]).astype(np.float32)),
                               synth.VIS_FRAME: torch.from_numpy(frames)}, out16_dtype=torch.bfloat16)
    sub = np.arange(0, V, 37)   # the oracle checks every 37th video (the full CPU pass would take ~30 s)
    ref, _ = O.frame_vis_net_forward({k: x[sub] for k, x in vin.items()}, frames[sub], synth.VIS_FRAME, sd, [synth.VIS_FRAME], 8)
    assert max_abs(emb32[torch.from_numpy(sub).cuda()], ref) <= 2e-6
    np.testing.assert_allclose(emb32.norm(dim=2).cpu().numpy(), 1.0, atol=1e-6)
    # retrieval over the whole gallery: queries = noisy copies of 1000 gallery embeddings
    Q = 1000
    gt = torch.arange(0, Q, device="cuda", dtype=torch.int32) * 11
    n = torch.randn(Q, 8, 512, generator=torch.Generator(device="cuda").manual_seed(4), device="cuda")
    q = emb32[gt.long()] + 6.0 * n / n.norm(dim=2, keepdim=True)
    q16 = (q / q.norm(dim=2, keepdim=True)).reshape(Q, -1).to(torch.bfloat16)
    res = GalleryIndex(emb16.reshape(V, -1), V, 8).search(q16, gt, 10)
    dense = ops.sim_dense(q16, emb16.reshape(V, -1), 0.125).cpu().numpy()
    np.testing.assert_array_equal(res.rank0.cpu().numpy(), O.tie_rule_rank(dense, gt.cpu().numpy()))
    np.testing.assert_array_equal(res.metrics.cpu().numpy()[:4], O.metrics_from_rank0(res.rank0.cpu().numpy())[:4])
