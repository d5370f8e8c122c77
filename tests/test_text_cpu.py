"""Text front-end, CPU: the oracle restatement and the host-side tokeniser / vocabulary code against
outputs of the unmodified reference (tests/golden/text, written by tests/golden/make_golden_text.py)."""
import json
import os

import numpy as np
import pytest

from laff_b200 import synth
from laff_b200 import text as T
from laff_b200.bigfile import BigFile
from oracle import laff_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
D = os.path.join(HERE, "golden", "text")
META = json.load(open(os.path.join(D, "meta.json")))
GOLD = np.load(os.path.join(D, "golden.npz"))


@pytest.fixture(autouse=True)
def stopwords():
    T.TextTool.set_stopwords(META["stopwords_used"])
    yield
    T.TextTool._stopwords = None


def test_tokenizer_matches_reference():
    for c, a, b in zip(META["captions"], META["tokens"], META["tokens_nsw"]):
        assert T.TextTool.tokenize(c) == a == O.tokenize(c)
        assert T.TextTool.tokenize(c, remove_stopword=True) == b == O.tokenize(c, remove_stopword=True, stopwords=META["stopwords_used"])


def test_stopword_removal_without_a_list_fails_loudly(monkeypatch):
    T.TextTool._stopwords = None
    monkeypatch.delenv("LAFF_STOPWORDS_EN", raising=False)
    with pytest.raises(T.LaffError):
        T.TextTool.tokenize("a dog", remove_stopword=True)


def test_reference_vocabulary_pickles_load_without_the_reference():
    bow = T.load_vocab(os.path.join(D, "vocab_bow_nsw.pkl"))
    gru = T.load_vocab(os.path.join(D, "vocab_gru.pkl"))
    assert isinstance(bow, T.Vocabulary) and [bow[i] for i in range(len(bow))] == META["bow_words"]
    assert [gru[i] for i in range(len(gru))] == META["gru_words"]
    assert bow.find("zebra") == -1 and gru("zebra") == gru("<unk>")
    with pytest.raises(Exception, match="word out of vocab"):
        bow("zebra")
    idx = T.IndexVec(os.path.join(D, "vocab_gru.pkl"))
    for c, ref in zip(META["captions"], META["index"]):
        assert idx.encoding(c).tolist() == ref
        assert O.index_encoding(O.tokenize(c), gru.word2idx).tolist() == ref


def test_oracle_bow_w2v_match_reference():
    bow = T.load_vocab(os.path.join(D, "vocab_bow_nsw.pkl"))
    w2v = BigFile(os.path.join(D, "w2v"))
    table = np.asarray(w2v.matrix())
    assert np.array_equal(table, GOLD["w2v_table"])
    for i, c in enumerate(META["captions"]):
        words = O.tokenize(c, remove_stopword=True, stopwords=META["stopwords_used"])
        assert np.array_equal(O.bow_encoding(words, bow.word2idx), GOLD["bow_enc"][i])
        assert np.array_equal(O.w2v_encoding(words, w2v.name2index, table), GOLD["w2v_enc"][i])     # float64, bit-exact
    assert np.array_equal(GOLD["bow_enc"].astype(np.float32), GOLD["bow_module"])
    assert np.array_equal(GOLD["w2v_enc"].astype(np.float32), GOLD["w2v_module"])
    assert not GOLD["w2v_enc"][2].any() and not GOLD["bow_enc"][4].any()                             # empty bags


def gru_params(tag, shapes):
    return {k: np.asarray(synth.param(71, "gru_%s/%s" % (tag, k), shp)) for k, shp in shapes.items()}


def test_oracle_gru_matches_reference():
    gru = T.load_vocab(os.path.join(D, "vocab_gru.pkl"))
    ids = [np.array(v) for v in META["index"]]
    sd = {k[len("gru_small/"):]: GOLD[k] for k in GOLD.files if k.startswith("gru_small/")}
    for pooling in ("mean", "last", "mean_last"):
        out = O.gru_encoder(ids, sd["we.weight"], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"],
                            sd["rnn.bias_hh_l0"], pooling)
        np.testing.assert_allclose(out, GOLD["gru_small_%s" % pooling], rtol=0, atol=2e-6)
    V = len(gru)
    p = gru_params("full", {"we.weight": (V, 500), "rnn.weight_ih_l0": (3072, 500), "rnn.weight_hh_l0": (3072, 1024),
                            "rnn.bias_ih_l0": (3072,), "rnn.bias_hh_l0": (3072,)})
    out = O.gru_encoder(ids, p["we.weight"], p["rnn.weight_ih_l0"], p["rnn.weight_hh_l0"], p["rnn.bias_ih_l0"], p["rnn.bias_hh_l0"])
    np.testing.assert_allclose(out, GOLD["gru_full_mean"], rtol=0, atol=5e-6)


def test_norm_other_than_zero_fails_like_the_reference():
    with pytest.raises(AttributeError):
        T.BowVec(os.path.join(D, "vocab_bow_nsw.pkl"), norm=2)


def test_native_batch_tokeniser_equals_the_python_path():
    """csrc/tokenize.cu (host code of the C ABI) against TextTool.tokenize + the dict lookups, on the golden captions and
    on awkward inputs: non-ASCII text, digits, empty and all-separator strings, very long words, repeated words."""
    bow = T.BowVecNSW(os.path.join(D, "vocab_bow_nsw.pkl"))
    w2v = T.W2VecNSW(os.path.join(D, "w2v"))
    idx = T.IndexVec(os.path.join(D, "vocab_gru.pkl"))
    caps = list(META["captions"]) + ["", "   ", "?!--", "Dog dog DOG d0g", "café naïve 狗 dog", "x" * 5000 + " guitar",
                                     "the-man_playing\tguitar\r\nand 42"]
    off, ids = bow.token_csr(caps)
    assert [ids[off[i]:off[i + 1]].tolist() for i in range(len(caps))] == [bow.token_ids(c) for c in caps]
    off, ids = w2v.word_csr(caps)
    assert [ids[off[i]:off[i + 1]].tolist() for i in range(len(caps))] == [w2v.word_ids(c) for c in caps]
    mat, lengths = idx.encoding_batch(caps)
    for i, c in enumerate(caps):
        ref = idx.encoding(c).tolist()
        assert lengths[i] == len(ref) and mat[i, : lengths[i]].tolist() == ref and not mat[i, lengths[i]:].any()
    for c, ref in zip(META["captions"], META["index"]):                   # and against the reference itself
        assert idx.encoding_batch([c])[0][0].tolist() == ref
    bow_plain = T.BowVec(None, vocab=T.load_vocab(os.path.join(D, "vocab_bow_nsw.pkl")))   # a non-'gru' vocabulary has no <unk>
    with pytest.raises(Exception, match="word out of vocab"):
        T.IndexVec(None, vocab=bow_plain.vocab).encoding_batch(["zebra"])


def test_sparse_rows_host_logic():
    """ops.SparseRows (CSR BoW batch): shape, exact CSR form of a count matrix, host slices carry only their ids."""
    import torch
    from laff_b200 import ops
    counts = torch.tensor([[0., 2., 0., 1.], [0., 0., 0., 0.], [3., 0., 0., 0.], [0., 1., 1., 1.]])
    x = ops.SparseRows.from_dense(counts)
    assert x.shape == (4, 4) and x.offsets.tolist() == [0, 3, 3, 6, 9] and x.ids.tolist() == [1, 1, 3, 0, 0, 0, 1, 2, 3]
    s = x[2:4]
    assert s.shape == (2, 4) and s.base == 3 and s.ids.tolist() == [0, 0, 0, 1, 2, 3] and s.offsets.tolist() == [3, 6, 9]
    assert x[1:2].ids.numel() == 0 and x[4:4].shape == (0, 4)
    assert s[1:2].base == 6 and s[1:2].ids.tolist() == [1, 2, 3]
    y = ops.SparseRows.from_lists([[1, 1, 3], [], [0, 0, 0], [1, 2, 3]], 4)
    assert y.offsets.tolist() == x.offsets.tolist() and y.ids.tolist() == x.ids.tolist()
    assert x.nbytes() == 5 * 8 + 9 * 4
    with pytest.raises(ops.LaffError):
        ops.SparseRows.from_dense(torch.tensor([[0.5]]))
    with pytest.raises(ops.LaffError):
        x[::2]
