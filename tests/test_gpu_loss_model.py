"""GPU parity of the margin ranking loss and of the model-level drop-in surface
(predict / compute_loss / get_model) through the C ABI.

Tolerances: loss relative error <= 1e-5, gradients <= 1e-4 relative to the largest gradient entry."""
import numpy as np
import pytest
import torch

from helpers import cuda, load_numpy_state, max_abs, sd_from_npz, small_dims
from laff_b200 import config as cfg
from laff_b200 import loss as L
from laff_b200 import model as M
from laff_b200 import ops, synth
from oracle import laff_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore():
    L.set_precision("bf16")
    yield
    L.set_precision("bf16")


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return max_abs(a, b) / max(float(np.abs(b).max()), 1e-30)


def test_loss_vs_reference_golden(golden):
    d = golden("loss.npz")
    txt, vis = cuda(d["txt"]), cuda(d["vis"])
    for mv in (1, 0):
        for direction in ("t2i", "i2t", "bidir"):
            for style in ("sum", "mean"):
                tag = "mv%d_%s_%s" % (mv, direction, style)
                crit = L.MarginRankingLoss(margin=0.2, measure="cosine", max_violation=bool(mv), cost_style=style,
                                           direction=direction)
                t = txt.clone().requires_grad_(True)
                v = vis.clone().requires_grad_(True)
                loss = crit(t, v)            # [B, H, d]: the sum over heads of model/model.py:857-858 in one call
                loss.backward()
                assert abs(loss.item() - float(d[tag + "/loss"])) <= 1e-5 * abs(float(d[tag + "/loss"])), tag
                assert rel(t.grad, d[tag + "/d_txt"]) <= 1e-4, tag
                assert rel(v.grad, d[tag + "/d_vis"]) <= 1e-4, tag
                # head by head, 2-D, exactly like the reference's loop
                tot = sum(crit(txt[:, h, :].contiguous(), vis[:, h, :].contiguous()) for h in range(txt.shape[1]))
                assert abs(tot.item() - float(d[tag + "/loss"])) <= 1e-5 * abs(float(d[tag + "/loss"])), tag
    for mv in (1, 0):
        for direction in ("t2i", "bidir"):
            tag = "score_mv%d_%s" % (mv, direction)
            crit = L.MarginRankingLossWithScore(margin=0.2, max_violation=bool(mv), cost_style="sum", direction=direction)
            s = cuda(d["score"]).requires_grad_(True)
            loss = crit(s)
            loss.backward()
            assert abs(loss.item() - float(d[tag + "/loss"])) <= 1e-5 * abs(float(d[tag + "/loss"])), tag
            assert max_abs(s.grad, d[tag + "/d_score"]) <= 1e-6, tag


@pytest.mark.parametrize("kind", ["random", "correlated"])
def test_training_step_loss_c3_vs_oracle(kind):
    """BASELINE config C3: B = 128, H = 8, d_h = 512, margin 0.2, t2i, max_violation, sum."""
    r = synth.rng_for(4, "c3" + kind)
    B, H, dh = 128, 8, 512
    vis = r.standard_normal((B, H, dh)).astype(np.float32)
    txt = r.standard_normal((B, H, dh)).astype(np.float32) if kind == "random" else \
        (vis + 3.0 * r.standard_normal((B, H, dh))).astype(np.float32)
    loss, d_txt, d_vis = ops.mrl_forward_backward(cuda(txt), cuda(vis), 0.2, True, "t2i", "sum")
    rl, rt, rv = O.multi_head_loss(txt.astype(np.float64), vis.astype(np.float64), 0.2, True, "sum", "t2i", want_grad=True)
    assert rl > 0
    assert abs(loss.item() - rl) <= 1e-5 * abs(rl)
    assert rel(d_txt, rt) <= 1e-4 and rel(d_vis, rv) <= 1e-4
    # odd batch size / single head
    l1, _, _ = ops.mrl_forward_backward(cuda(txt[:37, 0]), cuda(vis[:37, 0]), 0.2, True, "bidir", "mean")
    r1 = O.margin_ranking_loss(txt[:37, 0].astype(np.float64), vis[:37, 0].astype(np.float64), 0.2, True, "mean", "bidir")
    assert abs(l1.item() - r1) <= 1e-5 * abs(r1)


class FakeVisLoader:
    """Yields collate_vision-shaped dicts (data_provider.py:38-73)."""

    def __init__(self, feats, batch, frames=None):
        self.feats, self.batch, self.frames = feats, batch, frames
        self.n = next(iter(feats.values())).shape[0]
        self.dataset = type("D", (), {"length": self.n, "__len__": lambda s: self.n})()
        self.batch_size = batch

    def __iter__(self):
        for s in range(0, self.n, self.batch):
            e = min(self.n, s + self.batch)
            fd = {}
            if self.frames is not None:
                fd = {"mask_tensor": torch.ones(e - s, self.frames.shape[1]), synth.VIS_FRAME: torch.from_numpy(self.frames[s:e])}
            yield {"vis_feat_dict": {k: torch.from_numpy(v[s:e]) for k, v in self.feats.items()}, "idxs": list(range(s, e)),
                   "vis_ids": ["video%d" % i for i in range(s, e)], "vis_frame_feat_dict": fd, "vis_origin_frame_tuple": (None,)}


class FakeTxtLoader:
    """Yields collate_text-shaped tuples (data_provider.py:76-89)."""

    def __init__(self, feats, batch):
        self.feats, self.batch = feats, batch
        self.n = next(iter(feats.values())).shape[0]
        self.dataset = type("D", (), {"__len__": lambda s: self.n})()

    def __len__(self):
        return (self.n + self.batch - 1) // self.batch

    def __iter__(self):
        for s in range(0, self.n, self.batch):
            e = min(self.n, s + self.batch)
            yield ({k: torch.from_numpy(v[s:e]) for k, v in self.feats.items()}, list(range(s, e)),
                   ["video%d#0" % i for i in range(s, e)])


def test_predict_and_model_surface_vs_oracle(golden):
    """get_model('LAFF') + load_state_dict + predict(txt_loader, vis_loader) -> (scores, txt_ids, vis_ids), then the
    predictor's rank extraction, against the oracle pipeline on the same inputs (small golden parameters)."""
    d = golden("fusion_small.npz")
    D, H = int(d["meta"][0]), int(d["meta"][1])
    dims = small_dims(d)
    c = cfg.laff_config(D, H, dims)
    model = M.get_model("LAFF", "cuda", c)
    sd = {"vis_net." + k: v for k, v in sd_from_npz(d, "vsd/").items()}
    sd.update({"txt_net." + k: v for k, v in sd_from_npz(d, "tsd/").items()})
    sd["txt_net.encoder.rnn_encoder.we.weight"] = np.zeros((3, 3), np.float32)  # GRU encoder params pass through (strict=False)
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    assert not missing and unexpected == ["txt_net.encoder.rnn_encoder.we.weight"]
    n = 23
    names = [str(x) for x in d["vis_names"]]
    vis_in = {nm: synth.feature(3, "v/" + nm, n, dm, "dense" if nm == synth.VIS_CLIP_FT else "relu")
              for nm, dm in zip(names, [int(x) for x in d["vis_dims"]])}
    txt_in = {"gru": synth.feature(3, "t/gru", n, dims["gru"]), "bow": synth.feature(3, "t/bow", n, dims["bow"], "bow"),
              "w2v": synth.feature(3, "t/w2v", n, dims["w2v"]), "clip": synth.feature(3, "t/clip", n, dims["clip"])}
    L.set_precision("bf16x3")
    scores, txt_ids, vis_ids = model.predict(FakeTxtLoader(txt_in, 5), FakeVisLoader(vis_in, 7), "cosine", record_emb=True)
    assert scores.shape == (n, n) and scores.dtype == np.float32 and isinstance(scores, np.ndarray)
    assert txt_ids[3] == "video3#0" and vis_ids[22] == "video22"
    ov, _ = O.vis_net_forward(vis_in, sd_from_npz(d, "vsd/"), [synth.VIS_CLIP_FT], H)
    ot, _ = O.txt_net_forward(txt_in, sd_from_npz(d, "tsd/"), ["CLIP_encoder"], H)
    ref = O.txt2vis_matrix(ot, ov)
    assert max_abs(scores, ref) <= 2e-5
    # predictor.py:232-246 on our scores: same ranks / metrics as on the oracle's scores (no near ties in this case)
    gt = np.arange(n)
    np.testing.assert_array_equal(O.tie_rule_rank(scores, gt), O.tie_rule_rank(ref, gt))
    # second call reuses the cached gallery embeddings (record_emb, model/model.py:1026-1034)
    s2, _, _ = model.predict(FakeTxtLoader(txt_in, 23), FakeVisLoader(vis_in, 7), "cosine", record_emb=True)
    assert max_abs(s2, scores) <= 1e-7
    with pytest.raises(Exception, match="invalid"):
        model.predict(FakeTxtLoader(txt_in, 5), FakeVisLoader(vis_in, 7), "nope")
    # compute_loss on the embeddings (LAFF multi-space branch, model/model.py:2036-2038)
    t = model.txt_net(dict((k, torch.from_numpy(v)) for k, v in txt_in.items()))
    v = model.vis_net(dict((k, torch.from_numpy(x)) for k, x in vis_in.items()))
    loss, items = model.compute_loss(v, t)
    rl = O.multi_head_loss(ot, ov, 0.2, True, "sum", "t2i")
    assert abs(loss.item() - float(rl)) <= 2e-4 * max(1.0, abs(float(rl))) and "triplet_loss" in items
    with pytest.raises(ops.LaffError):      # a training step needs model.train() (tests/test_gpu_train.py covers the step)
        model({}, 0)
    with pytest.raises(AssertionError):
        M.get_model("W2VVPP", "cuda", c)


def test_frame_laff_model_predict(golden):
    d = golden("frame_small.npz")
    D, H = int(d["meta"][0]), int(d["meta"][1])
    names = [str(n) for n in d["names"]]
    dims = dict(zip(names, [int(x) for x in d["dims"]]))
    sm = {"clip": dims[synth.VIS_FRAME], "c3d": dims[synth.VIS_C3D], "tf": dims[synth.VIS_TF], "x3d": dims[synth.VIS_X3D],
          "ircsn": dims[synth.VIS_IRCSN], "gru": 40, "bow": 56, "w2v": 20}
    model = M.get_model("FrameLAFF", "cuda", cfg.frame_laff_config(D, H, sm))
    load_numpy_state(model.vis_net, sd_from_npz(d, "sd/"))
    L.set_precision("bf16x3")
    vis_in = {n: d["vin/" + n] for n in names if n != synth.VIS_FRAME}
    embs = []
    for out in FakeVisLoader(vis_in, 2, frames=d["frames"]):
        embs.append(model._encode_vis(out)[0])
    assert max_abs(torch.cat(embs, 0), d["emb"]) <= 2e-5


def test_predict_large_gallery_branch_returns_ranked_scores(golden, tmp_path):
    """predict() above LARGE_GALLERY videos takes the predict_batch branch like the reference (model/model.py:1020-1021)
    but hands back a retrieval.RankedScores instead of a dense host matrix: the predictor's evaluation and writers run
    on it with the fused sweep and give the dense path's ranks, metrics and files; np.asarray() still yields the
    matrix when it is small; vis loaders that deliver the videos in shuffled batches land in dataset order."""
    import pickle
    from laff_b200 import predictor as P
    from laff_b200.retrieval import RankedScores
    d = golden("fusion_small.npz")
    D, H = int(d["meta"][0]), int(d["meta"][1])
    dims = small_dims(d)
    c = cfg.laff_config(D, H, dims)
    model = M.get_model("LAFF", "cuda", c)
    sd = {"vis_net." + k: v for k, v in sd_from_npz(d, "vsd/").items()}
    sd.update({"txt_net." + k: v for k, v in sd_from_npz(d, "tsd/").items()})
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    n = 61
    names = [str(x) for x in d["vis_names"]]
    vis_in = {nm: synth.feature(3, "v/" + nm, n, dm, "dense" if nm == synth.VIS_CLIP_FT else "relu")
              for nm, dm in zip(names, [int(x) for x in d["vis_dims"]])}
    txt_in = {"gru": synth.feature(3, "t/gru", n, dims["gru"]), "bow": synth.feature(3, "t/bow", n, dims["bow"], "bow"),
              "w2v": synth.feature(3, "t/w2v", n, dims["w2v"]), "clip": synth.feature(3, "t/clip", n, dims["clip"])}
    dense, txt_ids, vis_ids = model.predict(FakeTxtLoader(txt_in, 9), FakeVisLoader(vis_in, 7), "cosine")
    assert isinstance(dense, np.ndarray)
    model.LARGE_GALLERY = 50
    try:
        pred, txt_ids2, vis_ids2 = model.predict(FakeTxtLoader(txt_in, 9), FakeVisLoader(vis_in, 7), "cosine")
    finally:
        del model.LARGE_GALLERY
    assert isinstance(pred, RankedScores) and pred.shape == (n, n) and txt_ids2 == txt_ids and vis_ids2 == vis_ids
    # the 16-bit embeddings of the fused kernel vs predict()'s renormalise-then-round: one 16-bit ulp on a few components
    assert np.abs(np.asarray(pred) - dense).max() <= 2e-5
    own = np.asarray(pred)
    gt = P.gt_index(txt_ids, vis_ids)
    res = pred.search(torch.from_numpy(gt), 10)
    np.testing.assert_array_equal(res.rank0.cpu().numpy(), O.tie_rule_rank(own, gt))
    np.testing.assert_array_equal(res.topk_idx.cpu().numpy(), O.tie_rule_topk(own, 10)[1])
    caps = {t: "caption of %s" % t for t in txt_ids}
    a = P.evaluate_and_write(pred, txt_ids, vis_ids, str(tmp_path / "ranked"), str(tmp_path / "res" / "r.txt"), "m", None, captions=caps)
    b = P.evaluate_and_write(own, txt_ids, vis_ids, str(tmp_path / "dense"), str(tmp_path / "res2" / "r.txt"), "m", None, captions=caps)
    np.testing.assert_allclose(a["t2v"], b["t2v"], atol=1e-9)
    da, db = pickle.load(open(tmp_path / "ranked" / "t2v.pkl", "rb")), pickle.load(open(tmp_path / "dense" / "t2v.pkl", "rb"))
    assert list(da) == list(db) and all(da[k]["rank_list"] == db[k]["rank_list"] for k in da)
    a2 = P.evaluate_and_write(pred, txt_ids, vis_ids, str(tmp_path / "adhoc"), None, "m", None, captions=caps, with_ground_truth=False)
    lines = open(a2["pred_result_file"]).read().splitlines()
    assert len(lines) == n and all(len(l.split()) == 1 + 2 * (n - 1) for l in lines)
    big = RankedScores(pred.index, pred.q16)
    big.dense_limit_bytes = 100
    with pytest.raises(MemoryError):
        big.numpy()


def test_mrl_with_score_accepts_a_strided_view():
    """MarginRankingLossWithScore on a score block sliced out of a wider matrix (row pitch > B): same loss and gradient
    as on the contiguous copy, and nothing is written outside the [B, B] gradient (the C entry point uses one pitch for
    the scores and their gradient)."""
    B = 37
    wide = torch.randn(B, 3 * B + 5, device="cuda")
    view = wide[:, 7:7 + B]
    assert not view.is_contiguous()
    l1, d1 = ops.mrl_score_forward_backward(view, 0.2, True, "t2i", "sum")
    l2, d2 = ops.mrl_score_forward_backward(view.contiguous(), 0.2, True, "t2i", "sum")
    assert d1.shape == (B, B) and d1.is_contiguous() and torch.equal(d1, d2) and float(l1) == float(l2)
    crit = L.MarginRankingLossWithScore(margin=0.2, max_violation=True, cost_style="sum", direction="bidir")
    v = view.clone().requires_grad_(True)
    w = wide.clone().requires_grad_(True)
    crit(v).backward()
    crit(w[:, 7:7 + B]).backward()
    assert torch.equal(w.grad[:, 7:7 + B], v.grad) and float(w.grad[:, :7].abs().sum()) == 0 and float(w.grad[:, 7 + B:].abs().sum()) == 0
