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
    assert torch.equal(from_strings, from_feats)
    np.testing.assert_allclose(from_strings.norm(dim=2).cpu().numpy(), 1.0, atol=1e-6)
