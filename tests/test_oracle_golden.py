"""Pin the CPU oracle (oracle/laff_oracle.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from laff_b200 import synth
from oracle import laff_oracle as O

EMB_TOL = 2e-6  # fp32 reference (ATen sgemm) vs fp32 numpy restatement: summation order only


def _sd(d, prefix):
    return {k[len(prefix):]: d[k] for k in d.files if k.startswith(prefix)}


def _fusion_inputs(d):
    names = [str(n) for n in d["vis_names"]]
    vis_in = {n: d["vin/" + n] for n in names}
    txt_in = {k: d["tin/" + k] for k in ("gru", "bow", "w2v", "clip")}
    return names, vis_in, txt_in, _sd(d, "vsd/"), _sd(d, "tsd/")


@pytest.mark.parametrize("fname", ["fusion_small.npz", "fusion_small_ave_mul.npz", "fusion_small_bf16in.npz"])
def test_fusion_small(golden, fname):
    d = golden(fname)
    D, H, rows, seed, with_ave, mul, _ = [int(x) for x in d["meta"]]
    names, vis_in, txt_in, vsd, tsd = _fusion_inputs(d)
    v, va = O.vis_net_forward(vis_in, vsd, [synth.VIS_CLIP_FT], H, bool(with_ave), bool(mul))
    t, ta = O.txt_net_forward(txt_in, tsd, ["CLIP_encoder"], H, bool(with_ave), bool(mul))
    assert v.shape == d["vis_emb"].shape == (rows, H, D // H)
    np.testing.assert_allclose(v, d["vis_emb"], atol=EMB_TOL, rtol=0)
    np.testing.assert_allclose(t, d["txt_emb"], atol=EMB_TOL, rtol=0)
    np.testing.assert_allclose(va, d["vis_att"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(ta, d["txt_att"], atol=1e-5, rtol=0)
    # every head is unit norm
    np.testing.assert_allclose(np.linalg.norm(v, axis=2), 1.0, atol=1e-5)


def regen_full(d):
    """Regenerate inputs and parameters of the full-dimension golden case from its seeds."""
    D, H, rows, seed, with_ave, mul, bf16_in = [int(x) for x in d["meta"]]
    names = [str(n) for n in d["vis_names"]]
    vdims = [int(x) for x in d["vis_dims"]]
    rnd = synth.bf16_round if bf16_in else (lambda a: a)
    vsd = {str(k): synth.param(seed, str(k), eval(str(s))) for k, s in zip(d["vsd_keys"], d["vsd_shapes"])}
    tsd = {str(k): synth.param(seed + 1, str(k), eval(str(s))) for k, s in zip(d["tsd_keys"], d["tsd_shapes"])}
    if bf16_in:
        for sd in (vsd, tsd):
            for k in sd:
                if k.endswith("fc1.weight"):
                    sd[k] = synth.bf16_round(sd[k])
    vis_in = {}
    for n, dim in zip(names, vdims):
        x = synth.feature(seed, "vis/" + n, rows, dim, "dense" if n == synth.VIS_CLIP_FT else "relu")
        vis_in[n] = x if n == synth.VIS_CLIP_FT else rnd(x)
    g, b, w, c = [int(x) for x in d["txt_dims"]]
    txt_in = {"gru": rnd(synth.feature(seed, "txt/gru", rows, g)), "bow": synth.feature(seed, "txt/bow", rows, b, "bow"),
              "w2v": rnd(synth.feature(seed, "txt/w2v", rows, w)), "clip": synth.feature(seed, "txt/clip", rows, c)}
    return H, vis_in, txt_in, vsd, tsd


def test_fusion_full_dims(golden):
    d = golden("fusion_full_bf16in.npz")
    H, vis_in, txt_in, vsd, tsd = regen_full(d)
    v, _ = O.vis_net_forward(vis_in, vsd, [synth.VIS_CLIP_FT], H)
    t, _ = O.txt_net_forward(txt_in, tsd, ["CLIP_encoder"], H)
    assert v.shape == (4, 8, 512)
    np.testing.assert_allclose(v, d["vis_emb"], atol=EMB_TOL, rtol=0)
    np.testing.assert_allclose(t, d["txt_emb"], atol=EMB_TOL, rtol=0)


@pytest.mark.parametrize("fname", ["frame_small.npz", "frame_small_ragged.npz"])
def test_frame_laff(golden, fname):
    d = golden(fname)
    D, H = int(d["meta"][0]), int(d["meta"][1])
    names = [str(n) for n in d["names"]]
    sd = _sd(d, "sd/")
    vis_in = {n: d["vin/" + n] for n in names if n != synth.VIS_FRAME}
    fe = O.frame_attention_forward(d["frames"], sd, synth.VIS_FRAME)
    np.testing.assert_allclose(fe, d["frame_emb"], atol=EMB_TOL, rtol=0)
    emb, _ = O.frame_vis_net_forward(vis_in, d["frames"], synth.VIS_FRAME, sd, [synth.VIS_FRAME], H)
    np.testing.assert_allclose(emb, d["emb"], atol=EMB_TOL, rtol=0)


def test_attention_variants(golden):
    d = golden("attention_variants.npz")
    Y = d["Y"]
    for with_ave in (0, 1):
        for mul in (0, 1):
            tag = "ave%d_mul%d" % (with_ave, mul)
            sd = _sd(d, tag + "/sd/")
            out, att = O.multi_head_attention(Y, sd, "", 8, bool(with_ave), bool(mul))
            np.testing.assert_allclose(out, d[tag + "/out"], atol=EMB_TOL, rtol=0)
            np.testing.assert_allclose(att, d[tag + "/att"], atol=1e-6, rtol=0)


def test_similarity_and_ranks(golden):
    d = golden("sim_eval.npz")
    Q, H = 64, 8
    for qk, gk, sk in (("q_bf16", "g_bf16", "scores_bf16"), ("q", "g", "scores_fp32")):
        s = O.txt2vis_matrix(d[qk].reshape(Q, H, -1), d[gk].reshape(Q, H, -1))
        np.testing.assert_allclose(s, d[sk], atol=2e-7, rtol=0)
    s = d["scores_bf16"]
    gt = np.arange(Q)
    # the reference's own argsort path restated
    np.testing.assert_array_equal(O.argsort_rank(s, gt), d["rank0"])
    # the documented tie rule equals the reference wherever the ground truth has no exact tie, and always lies in
    # the interval an (unstable) argsort can return; ties are planted at videos (3, 7) and (40, 41)
    lo, hi = O.rank_bounds(s, gt)
    tr = O.tie_rule_rank(s, gt)
    assert np.all((lo <= d["rank0"]) & (d["rank0"] <= hi))
    assert np.all((lo <= tr) & (tr <= hi))
    untied = lo == hi
    assert untied.sum() >= Q - 4 and (~untied).sum() >= 2
    np.testing.assert_array_equal(tr[untied], d["rank0"][untied])
    np.testing.assert_array_equal(tr, np.array([np.where(np.argsort(s[i], kind="stable")[::-1] == i)[0][0] for i in range(Q)]))
    tv, ti = O.tie_rule_topk(s, 10)
    ref_top = d["argsort"][:, ::-1][:, :10]
    np.testing.assert_array_equal(np.take_along_axis(s, ti, 1), np.take_along_axis(s, ref_top, 1))  # same scores
    np.testing.assert_allclose(O.eval_qry2retro(s, 1), d["eval_qry2retro"], rtol=1e-12)
    label = np.zeros_like(s)
    label[np.arange(Q), d["rank0"]] = 1
    np.testing.assert_allclose(O.eval_label_matrix(label), d["eval_label"], rtol=1e-12)
    np.testing.assert_allclose(O.cosine_sim_np(d["q"], d["g"]), d["np_cosine"], atol=1e-6)
    # mean-over-heads of unit-norm heads == one D-wide dot / H (the form the tensor-core kernel computes): exact on the
    # fp32 unit-norm embeddings up to summation order, and within bf16 operand rounding (2^-9 relative per component)
    # once the operands are rounded (the T2 tier: tensor-core operands vs the all-fp32 reference)
    np.testing.assert_allclose(O.mm_mean_heads(d["q"], d["g"], H), d["scores_fp32"], atol=3e-7)
    np.testing.assert_allclose(O.mm_mean_heads(d["q_bf16"], d["g_bf16"], H), d["scores_fp32"], atol=2e-3)


def test_reference_tie_order_is_not_the_stable_order(golden):
    """Documents why the tie rule is a stated convention: on the golden run numpy's default argsort ordered the two
    planted exact ties differently from argsort(kind='stable')."""
    d = golden("sim_eval.npz")
    s = d["scores_bf16"]
    assert s[3, 3] == s[3, 7] and s[41, 40] == s[41, 41]
    stable = np.argsort(s, axis=1, kind="stable")
    assert not np.array_equal(stable, d["argsort"])
    assert np.array_equal(np.take_along_axis(s, stable, 1), np.take_along_axis(s, d["argsort"], 1))


def test_l2norm(golden):
    d = golden("sim_eval.npz")
    x = d["l2_in"]
    np.testing.assert_allclose(O.l2norm(x), d["l2_torch"], atol=1e-7)
    np.testing.assert_allclose(O.l2norm(x, eps=0), d["l2_torch_eps0"], atol=1e-7)
    np.testing.assert_allclose(O.l2norm_np(x), d["l2_numpy"], atol=1e-7)
    assert np.all(O.l2norm(x)[4] == 0)  # all-zero row stays zero (norm = eps)


@pytest.mark.parametrize("name", ["odd", "even", "zeros", "big"])
def test_metrics_edge_cases(golden, name):
    d = golden("sim_eval.npz")
    rk = d["metrics_%s/rank0" % name]
    ref = d["metrics_%s/eval" % name]  # evaluation.eval: 1-based ranks -> (r1, r5, r10, medr, meanr, mir, mAP)
    r1, r5, r10, medr, meanr, mir = O.metrics_from_rank0(rk)
    np.testing.assert_allclose([r1, r5, r10, meanr, mir], [ref[0], ref[1], ref[2], ref[4], ref[5]], rtol=1e-12)
    # eval_qry2retro's floor(median(rank0)) + 1 and eval's floor(median(rank1)) agree except when the median is x.5
    assert medr in (ref[3], ref[3] + 1)
    assert medr == np.floor(np.median(rk)) + 1


def test_loss_forward_backward(golden):
    d = golden("loss.npz")
    txt, vis = d["txt"], d["vis"]
    for mv in (1, 0):
        for direction in ("t2i", "i2t", "bidir"):
            for style in ("sum", "mean"):
                tag = "mv%d_%s_%s" % (mv, direction, style)
                loss, dt, dv = O.multi_head_loss(txt, vis, 0.2, bool(mv), style, direction, want_grad=True)
                np.testing.assert_allclose(loss, d[tag + "/loss"], rtol=2e-6)
                np.testing.assert_allclose(dt, d[tag + "/d_txt"], atol=2e-6 * max(1.0, np.abs(d[tag + "/d_txt"]).max()))
                np.testing.assert_allclose(dv, d[tag + "/d_vis"], atol=2e-6 * max(1.0, np.abs(d[tag + "/d_vis"]).max()))
    for mv in (1, 0):
        for direction in ("t2i", "bidir"):
            tag = "score_mv%d_%s" % (mv, direction)
            loss, g = O.margin_ranking_loss_with_score(d["score"], 0.2, bool(mv), "sum", direction, want_grad=True)
            np.testing.assert_allclose(loss, d[tag + "/loss"], rtol=2e-6)
            np.testing.assert_allclose(g, d[tag + "/d_score"], atol=1e-6)


def test_compute_sim_errors():
    x = np.eye(3, dtype=np.float32)
    with pytest.raises(Exception, match="invalid"):
        O.compute_sim(x, x, "nope")
    with pytest.raises(Exception, match="Not implemented"):
        O.compute_sim(x, x, "euclidean")


def test_retrieve_cpu_matches_tie_rule():
    q, g, gt = synth.retrieval_embeddings(5, 40, 300, 8, 32, sigma=1.5)
    rank0, topk, m = O.retrieve_cpu(q, g, gt, 8, k=5, chunk=128, threads=2)
    s = O.txt2vis_matrix(q.reshape(40, 8, 32), g.reshape(300, 8, 32))
    np.testing.assert_array_equal(rank0, O.tie_rule_rank(s, gt))
    np.testing.assert_array_equal(topk, O.tie_rule_topk(s, 5)[1])
    assert m == O.metrics_from_rank0(rank0)


def test_dual_softmax_loss_oracle_matches_reference_autograd():
    """DualSoftmaxLoss (loss.py:291-310): value and autograd gradients of the unmodified reference (tests/golden/dsl.npz)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dsl.npz"))
    for tag in ("small", "b128"):
        txt, vis = d[tag + "/txt"], d[tag + "/vis"]
        total, gt_, gv_ = 0.0, np.zeros_like(txt, dtype=np.float64), np.zeros_like(vis, dtype=np.float64)
        for h in range(txt.shape[1]):
            l, a, b = O.dual_softmax_loss(txt[:, h], vis[:, h], 1000.0, want_grad=True)
            assert abs(l - d[tag + "/per_head"][h]) <= 1e-5 * abs(d[tag + "/per_head"][h])
            total += l
            gt_[:, h], gv_[:, h] = a, b
        assert abs(total - float(d[tag + "/loss"])) <= 1e-5 * abs(float(d[tag + "/loss"]))
        for got, key in ((gt_, "/d_txt"), (gv_, "/d_vis")):
            ref = d[tag + key]
            assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
        for temp in (1.0, 0.05):
            l, a, b = O.dual_softmax_loss(txt[:, 0], vis[:, 0], temp, want_grad=True)
            assert abs(l - float(d["%s/temp%g/loss" % (tag, temp)])) <= 1e-5 * abs(float(d["%s/temp%g/loss" % (tag, temp)]))
            assert np.abs(a - d["%s/temp%g/d_txt" % (tag, temp)]).max() <= 1e-4 * np.abs(d["%s/temp%g/d_txt" % (tag, temp)]).max()
            assert np.abs(b - d["%s/temp%g/d_vis" % (tag, temp)]).max() <= 1e-4 * np.abs(d["%s/temp%g/d_vis" % (tag, temp)]).max()
