"""T2 end-to-end parity on the reference-trained checkpoint (north_star: "top-k indices, ranks, R@K and
MedR identical"): tests/golden/trained_laff.npz is a checkpoint the UNMODIFIED reference trained on CPU together with
the reference's own ranks / top-10 / R@K / MedR on held-out C1- and C2-sized collections
(tests/golden/make_golden_trained.py).  The checkpoint goes through get_model + load_state_dict like predictor.py:160-167,
raw fp32 features go in, and the whole device pipeline (fused projection + LAFF pooling kernel -> similarity sweep with
rank + top-10 -> metrics) is compared with what the reference returned:

  * 'bf16x3' (3-term split operands, fp32-grade products): rank of the ground truth, top-10 lists, R@1/5/10 and MedR
    IDENTICAL, except queries where two of the reference's own fp32 scores lie within the accumulation noise of each
    other (enumerated from the fixture's margins, bounded, and their ranks within the number of such neighbours);
  * 'fp16' / 'bf16' operands (the tensor-core rate the benchmark runs at): fraction of queries whose rank moved,
    |dR@K| and dMedR reported and bounded -- the evidence behind the choice of the default operand type.
"""
import numpy as np
import pytest
import torch

from laff_b200 import config as cfg
from laff_b200 import loss as L
from laff_b200 import model as M
from laff_b200 import ops, synth
from laff_b200.retrieval import GalleryIndex
from test_trained_fixture_cpu import load_trained

pytestmark = pytest.mark.gpu
WINDOW = 1e-5     # fp32 reference noise (~1e-6) + residual of the 3-term bf16 split (8e-6)


@pytest.fixture(autouse=True)
def _restore():
    yield
    L.set_precision("bf16")


def build_model(g, sd):
    D, H = int(g["meta"][0]), int(g["meta"][1])
    c = cfg.laff_config(D, H, synth.TRAINED_DIMS)
    model = M.get_model("LAFF", torch.device("cuda"), c)
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    assert not unexpected and not [k for k in missing if "encoder." not in k], (missing, unexpected)
    return model.eval(), H


def run(model, H, vis, txt, precision):
    n = next(iter(vis.values())).shape[0]
    vin = {k: torch.from_numpy(x) for k, x in vis.items()}
    tin = {k: torch.from_numpy(x) for k, x in txt.items()}
    gt = torch.arange(n, device="cuda", dtype=torch.int32)
    if precision == "bf16x3":
        v32, _ = model.vis_net.encode(vin, precision="bf16x3")
        t32, _ = model.txt_net.encode(tin, precision="bf16x3")
        q, gal = ops.split3_16(t32.reshape(n, -1), 0), ops.split3_16(v32.reshape(n, -1), 1)
    else:
        dt = torch.float16 if precision == "fp16" else torch.bfloat16
        _, v16 = model.vis_net.encode(vin, out16_dtype=dt, precision=precision)
        _, t16 = model.txt_net.encode(tin, out16_dtype=dt, precision=precision)
        q, gal = t16.reshape(n, -1), v16.reshape(n, -1)
    res = GalleryIndex(gal, n, H).search(q, gt, 10)
    return res.rank0.cpu().numpy(), res.topk_idx.cpu().numpy(), res.topk_val.cpu().numpy(), res.metrics.cpu().numpy()


@pytest.mark.parametrize("tag", ["c1", "c2"])
def test_trained_checkpoint_ranks_identical_to_reference(tag):
    g, sd, noise = load_trained()
    model, H = build_model(g, sd)
    n = int(g[tag + "/n"])
    vis, txt = synth.latent_collection(int(g[tag + "/seed"]), n, **noise)
    ref_rank, ref_top, ref_m = g[tag + "/rank0"], g[tag + "/top10"], g[tag + "/metrics"]
    # ---- fp32-grade pipeline: identical
    r, ti, tv, m = run(model, H, vis, txt, "bf16x3")
    clean = g[tag + "/gt_gap"] > WINDOW
    assert clean.mean() > 0.97, clean.mean()
    np.testing.assert_array_equal(r[clean], ref_rank[clean])
    assert np.abs(r[~clean] - ref_rank[~clean]).max(initial=0) <= 2
    lists_clean = g[tag + "/min_gap_top11"] > WINDOW
    assert lists_clean.mean() > 0.85, lists_clean.mean()
    np.testing.assert_array_equal(ti[lists_clean], ref_top[lists_clean])
    assert np.abs(tv - g[tag + "/top_scores"][:, :10]).max() <= WINDOW
    flips = int((r != ref_rank).sum())
    assert m[3] == ref_m[3], (m[3], ref_m[3])                                      # MedR identical
    for i in range(3):                                                             # R@1/5/10: identical up to the enumerated near-ties
        assert abs(m[i] - ref_m[i]) <= 100.0 * flips / n + 1e-9, (i, m[i], ref_m[i], flips)
    report = {"bf16x3": (flips / n, [abs(m[i] - ref_m[i]) for i in range(3)], m[3] - ref_m[3])}
    # ---- 16-bit operands: reported and bounded
    for precision, bound_moved, bound_rk in (("fp16", 0.015, 0.11), ("bf16", 0.08, 0.21)):
        r16, ti16, tv16, m16 = run(model, H, vis, txt, precision)
        moved = float((r16 != ref_rank).mean())
        d = [abs(m16[i] - ref_m[i]) for i in range(3)]
        report[precision] = (moved, d, m16[3] - ref_m[3])
        assert moved <= bound_moved and max(d) <= bound_rk and abs(m16[3] - ref_m[3]) <= 1, (precision, moved, d, m16[3], ref_m[3])
        assert np.abs(tv16 - g[tag + "/top_scores"][:, :10]).max() <= (4e-4 if precision == "fp16" else 3e-3)
    print("\n%s trained fixture (reference R@1/5/10/MedR %s): " % (tag, np.round(ref_m[:4], 2)) +
          "; ".join("%s ranks moved %.2f%% max|dR@K| %.2f dMedR %+.0f" % (p, 100 * v[0], max(v[1]), v[2]) for p, v in report.items()))
    assert report["fp16"][0] <= report["bf16"][0] + 1e-9      # fp16 rounds 8x finer than bf16 at the same MMA rate


def test_full_dims_model_trained_on_device_ranks_like_the_oracle():
    """T2 at the shipped dimensions (D = 4096, 8 heads, gru 1024 / bow 3981 / w2v 500 / tf 768 / x3d, ircsn 2048): a LAFF
    model is trained here for 300 steps by the device training step (itself pinned to the reference's, tests/test_gpu_train.py)
    on the latent-factor collection, then 1000 held-out queries are ranked against 1000 held-out videos.  The oracle --
    pinned to the reference on the reference-trained fixture -- runs the same weights in numpy; the fp32-grade device
    pipeline must return its ranks / top-10 / R@K / MedR wherever the oracle's own margins exceed the noise window, and
    the 16-bit pipelines are reported and bounded like on the fixture."""
    from laff_b200 import ops as _ops
    from oracle import laff_oracle as O
    H, D, n_train, n_eval = 8, 4096, 4096, 1000
    noise = dict(vis_noise=2.0, cap_noise=1.3, txt_noise=2.0)     # wide features average their noise: more of it than on the fixture
    c = cfg.laff_config(D, H, synth.DIMS)
    c.dropout = 0.2
    torch.manual_seed(0)
    model = M.get_model("LAFF", torch.device("cuda"), c)
    vis_tr, txt_tr = synth.latent_collection(500, n_train, dims=synth.DIMS, **noise)
    pick = np.random.RandomState(3)
    model.train()
    losses = []
    for step in range(300):
        b = pick.choice(n_train, 128, replace=False)
        td = {"vis_feats": {k: torch.from_numpy(v[b]) for k, v in vis_tr.items()}, "captions": {k: torch.from_numpy(v[b]) for k, v in txt_tr.items()},
              "captions_task2": None, "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}
        losses.append(float(model(td, epoch=0)["triplet_loss"]))
    assert losses[-1] < 0.7 * losses[0], (losses[0], losses[-1])
    model.eval()
    vis, txt = synth.latent_collection(501, n_eval, dims=synth.DIMS, **noise)
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    vsd = {k[len("vis_net."):]: v for k, v in sd.items() if k.startswith("vis_net.")}
    tsd = {k[len("txt_net."):]: v for k, v in sd.items() if k.startswith("txt_net.")}
    ov, _ = O.vis_net_forward(vis, vsd, [synth.VIS_CLIP_FT], H)
    ot, _ = O.txt_net_forward(txt, tsd, ["CLIP_encoder"], H)
    s_ref = O.txt2vis_matrix(ot, ov)
    gt = np.arange(n_eval)
    r_ref = O.tie_rule_rank(s_ref, gt)
    m_ref = O.metrics_from_rank0(r_ref)
    assert 10 < m_ref[0] < 96, m_ref                      # a trained, non-saturated task
    r, ti, tv, m = run(model, H, vis, txt, "bf16x3")
    sg = s_ref[gt, gt][:, None]
    d = np.abs(s_ref - sg)
    d[gt, gt] = np.inf
    clean = d.min(1) > 2e-5                                # fp32 oracle noise + split residual at D = 4096
    assert clean.mean() > 0.95, clean.mean()
    np.testing.assert_array_equal(r[clean], r_ref[clean])
    assert np.abs(r[~clean] - r_ref[~clean]).max(initial=0) <= 3
    assert m[3] == m_ref[3] and all(abs(m[i] - m_ref[i]) <= 100.0 * (~clean).sum() / n_eval + 1e-9 for i in range(3))
    srt = -np.sort(-s_ref, axis=1)[:, :12]
    lists_clean = (srt[:, :-1] - srt[:, 1:])[:, :11].min(1) > 2e-5
    np.testing.assert_array_equal(ti[lists_clean], O.tie_rule_topk(s_ref, 10)[1][lists_clean])
    assert np.abs(tv - srt[:, :10]).max() <= 2e-5
    report = {}
    for precision, bound in (("fp16", 0.03), ("bf16", 0.15)):
        r16, _, tv16, m16 = run(model, H, vis, txt, precision)
        moved = float((r16 != r_ref).mean())
        report[precision] = (moved, max(abs(m16[i] - m_ref[i]) for i in range(3)), m16[3] - m_ref[3])
        assert moved <= bound and report[precision][1] <= 0.5 and abs(report[precision][2]) <= 1, (precision, report[precision])
    print("\nfull-dims model trained on the device (oracle R@1/5/10/MedR %s): bf16x3 ranks moved %.2f%%; " % (
        np.round(m_ref[:4], 2), 100 * float((r != r_ref).mean())) +
        "; ".join("%s ranks moved %.2f%% max|dR@K| %.2f dMedR %+.0f" % (p, 100 * v[0], v[1], v[2]) for p, v in report.items()))
    assert report["fp16"][0] <= report["bf16"][0] + 1e-9
