"""Text front-end on the GPU: laff_bow_counts / laff_gather_mean / the GRU encoder against the golden
outputs of the unmodified reference (tests/golden/text) and the oracle; and the text net driven from caption strings.

Bars: BoW counts and word-vector means bit-exact (integer counts; fp64 accumulation rounded once); GRU features
within 2e-5 of the reference's fp32 torch GRU (3-term bf16-split tensor-core products, T <= 21 recurrent steps)."""
import json
import os
import types

import numpy as np
import pytest
import torch

from helpers import load_numpy_state
from laff_b200 import config as cfg
from laff_b200 import model as M
from laff_b200 import synth
from laff_b200 import text as T
from oracle import laff_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
D = os.path.join(HERE, "golden", "text")
META = json.load(open(os.path.join(D, "meta.json")))
GOLD = np.load(os.path.join(D, "golden.npz"))
CAPS = META["captions"]


@pytest.fixture(autouse=True)
def stopwords():
    T.TextTool.set_stopwords(META["stopwords_used"])
    yield
    T.TextTool._stopwords = None


def t2v_objects():
    bow = T.BowVecNSW(os.path.join(D, "vocab_bow_nsw.pkl"))
    w2v = T.W2VecNSW(os.path.join(D, "w2v"))
    idx = T.IndexVec(os.path.join(D, "vocab_gru.pkl"))
    return bow, w2v, idx


def test_bow_and_w2v_encoders_bit_exact():
    bow, w2v, idx = t2v_objects()
    opt = types.SimpleNamespace(t2v_bow=bow, t2v_w2v=w2v)
    b = M.BoWTxtEncoder(opt)({"caption": CAPS})["text_features"]
    w = M.W2VTxtEncoder(opt)({"caption": CAPS})["text_features"]
    assert b.is_cuda and w.is_cuda
    assert np.array_equal(b.cpu().numpy(), GOLD["bow_module"])
    assert np.array_equal(w.cpu().numpy(), GOLD["w2v_module"])
    assert np.array_equal(bow.encoding(CAPS[3]), GOLD["bow_enc"][3])               # single-query API, same kernels
    assert np.array_equal(w2v.encoding(CAPS[0]).astype(np.float32), GOLD["w2v_module"][0])
    assert b.shape == (len(CAPS), bow.ndims) and M.BoWTxtEncoder(opt)({"caption": []})["text_features"].shape == (0, bow.ndims)


@pytest.mark.parametrize("tag,we_dim,H,poolings", [("small", 12, 32, ("mean", "last", "mean_last")), ("full", 500, 1024, ("mean",))])
def test_gru_encoder_vs_reference(tag, we_dim, H, poolings):
    _, _, idx = t2v_objects()
    for pooling in poolings:
        opt = types.SimpleNamespace(t2v_idx=idx, rnn_layer=1, we_dim=we_dim, rnn_size=H, pooling=pooling, we=None)
        enc = M.GruTxtEncoder(opt)
        sd = {k: np.asarray(synth.param(71, "gru_%s/%s" % (tag, k), tuple(v.shape))) for k, v in enc.state_dict().items()}
        load_numpy_state(enc, sd)
        enc = enc.cuda().eval()
        out = enc({"caption": CAPS})["text_features"].cpu().numpy()
        ref = GOLD["gru_%s_%s" % (tag, pooling)]
        assert out.shape == ref.shape
        assert np.abs(out - ref).max() <= 2e-5, (tag, pooling, np.abs(out - ref).max())
        ids = [np.array(v) for v in META["index"]]
        orc = O.gru_encoder(ids, sd["we.weight"], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"],
                            sd["rnn.bias_hh_l0"], pooling)
        assert np.abs(out - orc).max() <= 2e-5
    with pytest.raises(NotImplementedError):
        enc.train()({"caption": CAPS})


def test_text_net_from_caption_strings_equals_precomputed_features():
    """MultiScaleTxtEncoderAttention fed with strings (front-end encoders built from the config's vocabulary objects)
    == the same net fed with the per-encoder features computed separately; state_dict carries the reference's
    encoder.rnn_encoder.* keys."""
    bow, w2v, idx = t2v_objects()
    dims = dict(synth.DIMS)
    dims.update(bow=bow.ndims, w2v=w2v.ndims)
    c = cfg.laff_config(4096, 8, dims)
    c.t2v_bow, c.t2v_w2v, c.t2v_idx = bow, w2v, idx
    c.we_dim, c.rnn_size, c.rnn_layer, c.we = 500, 1024, 1, None
    net = M.MultiScaleTxtEncoderAttention(c)
    keys = set(net.state_dict().keys())
    assert {"encoder.rnn_encoder.we.weight", "encoder.rnn_encoder.rnn.weight_ih_l0", "encoder.rnn_encoder.rnn.weight_hh_l0",
            "encoder.rnn_encoder.rnn.bias_ih_l0", "encoder.rnn_encoder.rnn.bias_hh_l0"} <= keys
    load_numpy_state(net, {k: np.asarray(synth.param(9, k, tuple(v.shape))) for k, v in net.state_dict().items()})
    net = net.cuda().eval()
    clip = torch.from_numpy(np.random.RandomState(0).standard_normal((len(CAPS), 512)).astype(np.float32))
    from_strings, _ = net.encode({"caption": CAPS, "CLIP_encoding": clip})
    enc = dict(net.encoder.named_children())
    feats = {"gru": enc["rnn_encoder"]({"caption": CAPS})["text_features"], "bow": enc["bow_encoder"]({"caption": CAPS})["text_features"],
             "w2v": enc["w2v_encoder"]({"caption": CAPS})["text_features"], "clip": clip}
    from_feats, _ = net.encode(feats)
    # strings take the sparse BoW route (CSR ids, fp32 gather-sum over W^T); the dense count matrix goes through the
    # 16-bit tensor-core GEMM, which rounds W_bow: equal within that rounding, and bit-equal once both are sparse
    assert float((from_strings - from_feats).abs().max()) <= 6e-3
    sparse = dict(feats)
    sparse["bow_csr"] = enc["bow_encoder"]({"caption": CAPS}, sparse=True)["text_features"]
    del sparse["bow"]
    assert torch.equal(from_strings, net.encode(sparse)[0])
    np.testing.assert_allclose(from_strings.norm(dim=2).cpu().numpy(), 1.0, atol=1e-6)


# ---- sparse BoW projection ("BoW as sparse gather-sum of W_bow columns instead of a 3981-wide dense GEMM") ----
def _random_csr(seed, rows, vocab, max_len=12):
    r = synth.rng_for(seed, "csr")
    lists = []
    for i in range(rows):
        n = 0 if i % 17 == 5 else int(r.randint(1, max_len + 1))          # some captions have no known word
        ids = r.randint(0, vocab, size=n)
        if n >= 3 and i % 3 == 0:
            ids[1] = ids[0]                                                  # a repeated word counts twice
        lists.append(list(ids))
    lists[0] = list(r.randint(0, vocab, size=300))                          # a long caption: more than one 256-token pass
    lists[1] = [-1, 3, vocab + 7, 3]                                         # out-of-vocabulary markers are skipped
    return lists


@pytest.mark.parametrize("act,bn", [("tanh", False), ("relu", True), (None, False), ("sigmoid", True)])
def test_bow_project_sparse_equals_dense_reference(act, bn):
    """laff_bow_project against the reference's arithmetic on the dense count vector (txt2vec.py:56-63 + TransformNet,
    model/model.py:257-276) in float64."""
    from laff_b200 import ops
    rows, vocab, Dm = 70, 3981, 4096
    lists = _random_csr(5, rows, vocab)
    x = ops.SparseRows.from_lists(lists, vocab)
    counts = np.zeros((rows, vocab))
    for i, l in enumerate(lists):
        for t in l:
            if 0 <= t < vocab:
                counts[i, t] += 1
    r = synth.rng_for(6, "w")
    W = (r.standard_normal((Dm, vocab)) * 0.05).astype(np.float32)
    b = (r.standard_normal(Dm) * 0.1).astype(np.float32)
    sc, sh = (r.uniform(0.5, 1.5, Dm).astype(np.float32), r.standard_normal(Dm).astype(np.float32)) if bn else (None, None)
    z = counts @ W.astype(np.float64).T + b
    ref = {"tanh": np.tanh, "relu": lambda v: np.maximum(v, 0), "sigmoid": lambda v: 1 / (1 + np.exp(-v)), None: lambda v: v}[act](z)
    if bn:
        ref = ref * sc + sh
    wt = torch.from_numpy(np.ascontiguousarray(W.T)).cuda()
    cu = lambda a: None if a is None else torch.from_numpy(a).cuda()
    y = ops.bow_project(x.to("cuda"), wt, cu(b), act, cu(sc), cu(sh))
    assert np.abs(y.cpu().numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    assert np.array_equal(x.to("cuda").dense().cpu().numpy(), counts.astype(np.float32))      # CSR -> the reference's count vectors
    # row slices: a host slice carries only its own ids (base offset), a device slice keeps the id array
    for lo, hi in ((0, 1), (1, 40), (33, 70), (5, 5)):
        ys = ops.bow_project(x[lo:hi].to("cuda"), wt, cu(b), act, cu(sc), cu(sh))
        yd = ops.bow_project(x.to("cuda")[lo:hi], wt, cu(b), act, cu(sc), cu(sh))
        assert torch.equal(ys, y[lo:hi]) and torch.equal(yd, y[lo:hi])
        assert x[lo:hi].ids.numel() == sum(len(l) for l in lists[lo:hi])
    assert torch.equal(ops.SparseRows.from_dense(torch.from_numpy(counts)).to("cuda").dense().cpu(), torch.from_numpy(counts).float())


@pytest.mark.parametrize("single_kernel", [True, False])
def test_txt_net_sparse_bow_matches_oracle_and_dense_path(single_kernel):
    """The text net at the shipped dims fed with CSR BoW ids ('bow_csr') instead of the dense count matrix: fused
    embedding within 2e-6 of the oracle (T1: bf16-representable dense inputs; the sparse feature is projected in fp32),
    i.e. at least as close as the dense tensor-core path; string captions take the same route."""
    from laff_b200 import ops
    Q, H = 200, 8
    c = cfg.laff_config(4096, H, synth.DIMS)
    txt = M.MultiScaleTxtEncoderAttention(c)
    sd = {n: synth.bf16_round(synth.param(7, n, tuple(v.shape))) if n.endswith("fc1.weight") else synth.param(7, n, tuple(v.shape))
          for n, v in txt.state_dict().items()}
    load_numpy_state(txt, sd)
    txt = txt.cuda().eval()
    feats = {"gru": synth.bf16_round(synth.feature(7, "gru", Q, synth.DIMS["gru"])), "bow": synth.feature(7, "bow", Q, synth.DIMS["bow"], "bow"),
             "w2v": synth.bf16_round(synth.feature(7, "w2v", Q, synth.DIMS["w2v"])), "clip": synth.feature(7, "clip", Q, synth.DIMS["clip"])}
    ref, _ = O.txt_net_forward(feats, sd, ["CLIP_encoder"], H)
    M.set_single_kernel_fusion(single_kernel)
    try:
        dense, _ = txt.encode({n: torch.from_numpy(x) for n, x in feats.items()})
        sparse_in = {n: torch.from_numpy(x) for n, x in feats.items() if n != "bow"}
        sparse_in["bow_csr"] = ops.SparseRows.from_dense(torch.from_numpy(feats["bow"]))
        emb, emb16 = txt.encode(sparse_in, out16_dtype=torch.float16)
        pair_in = dict(sparse_in, bow_csr=(sparse_in["bow_csr"].offsets, sparse_in["bow_csr"].ids))
        emb_pair, _ = txt.encode(pair_in)
    finally:
        M.set_single_kernel_fusion(True)
    e_sparse, e_dense = np.abs(emb.cpu().numpy() - ref).max(), np.abs(dense.cpu().numpy() - ref).max()
    assert e_sparse <= 2e-6 and e_dense <= 2e-6, (e_sparse, e_dense)
    assert torch.equal(emb, emb_pair)
    assert np.abs(emb16.float().cpu().numpy() - ref).max() <= 5e-4
    txt.train()                                                   # the training step still takes the dense matrix
    assert txt._feature(sparse_in, "bow_encoder").shape == (Q, synth.DIMS["bow"])


def test_text_net_from_strings_uses_sparse_bow():
    """Caption strings in eval mode: the BoW front-end hands CSR ids to the projection; same embedding as the dense
    count vectors of the reference's BoWTxtEncoder."""
    from laff_b200 import ops
    bow, w2v, idx = t2v_objects()
    sp = bow.encode_sparse(CAPS)
    assert isinstance(sp, ops.SparseRows) and np.array_equal(sp.dense().cpu().numpy(), GOLD["bow_module"])
