"""Text front-end of the reference: tokeniser, vocabularies and the bag-of-words / word2vec / GRU
sentence encoders (textlib.py:27-112, txt2vec.py:12-166, model/model.py:322-434).

Strings are tokenised on the host (a regex and two dictionary lookups per word); everything numeric — the count
vectors, the word-vector means, the embedding lookup and the GRU recurrence — runs on the device through the C ABI
(`csrc/text.cu`, and the tcgen05 GEMM engine for the GRU's two matrix products).  There is no CPU arithmetic path:
`encoding()` of a single query runs the same device kernels with a batch of one.

The reference's stop-word list (`stopwords_en.txt`) is data of the reference tree and is not shipped here: point
`TextTool.set_stopwords()` (or the environment variable LAFF_STOPWORDS_EN) at it; asking for stop-word removal without a
list raises instead of silently keeping the words.
"""
from __future__ import annotations

import io
import os
import pickle
import re
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from ._capi import LaffError
from .bigfile import BigFile


# ----------------------------------------------------------------------------------------------------------------
# tokeniser and vocabulary (host)
# ----------------------------------------------------------------------------------------------------------------
class TextTool:
    """textlib.py:27-59 (English branch; the Chinese branch is Python-2 code in the reference and raises there)."""
    _stopwords: Optional[frozenset] = None

    @classmethod
    def set_stopwords(cls, words_or_path) -> None:
        if isinstance(words_or_path, (str, os.PathLike)):
            with open(words_or_path) as f:
                cls._stopwords = frozenset(map(str.strip, f.readlines()))
        else:
            cls._stopwords = frozenset(words_or_path)

    @classmethod
    def stopwords(cls) -> frozenset:
        if cls._stopwords is None:
            path = os.environ.get("LAFF_STOPWORDS_EN")
            if not path:
                raise LaffError("stop-word removal requested but no list is configured: call TextTool.set_stopwords(path) "
                                "or set LAFF_STOPWORDS_EN (the reference ships stopwords_en.txt)")
            cls.set_stopwords(path)
        return cls._stopwords

    @staticmethod
    def tokenize(input_str: str, clean: bool = True, language: str = "en", remove_stopword: bool = False) -> List[str]:
        if language != "en":
            raise NotImplementedError("only the English tokeniser of the reference is usable under Python 3")
        sent = input_str
        if clean:
            sent = sent.replace("\r", " ")
            sent = re.sub(r"[^A-Za-z0-9]", " ", sent).strip().lower()
        tokens = sent.split()
        if remove_stopword:
            stop = TextTool.stopwords()
            tokens = [x for x in tokens if x not in stop]
        return tokens


class Vocabulary:
    """textlib.py:81-112."""

    def __init__(self, encoding):
        self.word2idx = {}
        self.idx2word = {}
        self.encoding = encoding

    def add(self, word):
        if word not in self.word2idx:
            idx = len(self.word2idx)
            self.word2idx[word] = idx
            self.idx2word[idx] = word

    def find(self, word):
        return self.word2idx.get(word, -1)

    def __getitem__(self, index):
        return self.idx2word[index]

    def __call__(self, word):
        if word not in self.word2idx:
            if "gru" in self.encoding:
                return self.word2idx["<unk>"]
            raise Exception("word out of vocab: %s" % word)
        return self.word2idx[word]

    def __len__(self):
        return len(self.word2idx)


class _VocabUnpickler(pickle.Unpickler):
    """Reads the reference's vocabulary pickles (instances of textlib.Vocabulary) without the reference tree."""

    def find_class(self, module, name):
        if name == "Vocabulary" and module.split(".")[-1] in ("textlib", "text"):
            return Vocabulary
        return super().find_class(module, name)


def load_vocab(path_or_bytes) -> Vocabulary:
    if isinstance(path_or_bytes, (bytes, bytearray)):
        return _VocabUnpickler(io.BytesIO(path_or_bytes)).load()
    with open(path_or_bytes, "rb") as f:
        return _VocabUnpickler(f).load()


def _csr(lists: Sequence[Sequence[int]], device):
    offsets = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum([len(l) for l in lists], out=offsets[1:])
    flat = np.fromiter((i for l in lists for i in l), dtype=np.int32, count=int(offsets[-1]))
    return torch.from_numpy(offsets).to(device), torch.from_numpy(flat).to(device)


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise LaffError("laff_b200.text needs a CUDA device (no CPU fallback)")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


# ----------------------------------------------------------------------------------------------------------------
# native batch tokenisation + vocabulary lookup (csrc/tokenize.cu; host code of the C ABI, no GPU involved)
# ----------------------------------------------------------------------------------------------------------------
def _blob(strings: Sequence[str]):
    enc = [s.encode("utf-8") for s in strings]
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in enc], out=offsets[1:])
    return b"".join(enc), offsets


class NativeVocab:
    """A word -> id table inside the C library (open-addressing hash); also used as the stop-word set."""

    def __init__(self, words: Sequence[str], ids: Optional[Sequence[int]] = None):
        from . import _capi
        self._lib = _capi.lib()
        blob, offsets = _blob(words)
        idarr = None if ids is None else np.ascontiguousarray(ids, dtype=np.int32)
        self._keep = (blob, offsets, idarr)
        self.handle = self._lib.laff_vocab_create(blob, offsets.ctypes.data, None if idarr is None else idarr.ctypes.data, len(words))
        if not self.handle:
            raise LaffError("laff_vocab_create failed: %s" % (self._lib.laff_last_error() or b"?").decode())

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self._lib.laff_vocab_destroy(h)
            except Exception:
                pass


_STOP_NATIVE = {}


def _native_stopwords() -> NativeVocab:
    stop = TextTool.stopwords()
    nv = _STOP_NATIVE.get(id(stop))
    if nv is None or nv[0] is not stop:
        _STOP_NATIVE.clear()
        nv = _STOP_NATIVE[id(stop)] = (stop, NativeVocab(sorted(stop)))
    return nv[1]


def tokenize_lookup(captions: Sequence[str], vocab: NativeVocab, mode: int, remove_stopword: bool = False, unk_id: int = -1,
                    start_id: int = -1, end_id: int = -1):
    """TextTool.tokenize(clean=True) + vocabulary lookup for a batch, natively (laff_tokenize_lookup).
    mode 0: every token (unknown -> unk_id) between start_id / end_id; 1: known tokens in order; 2: distinct known ids,
    ascending.  Returns CSR (offsets int64 [n + 1], ids int32)."""
    from . import _capi
    lib = _capi.lib()
    blob, offsets = _blob(captions)
    stop = _native_stopwords().handle if remove_stopword else None
    cap = len(blob) // 2 + 3 * len(captions) + 8          # an upper bound on the number of tokens (+ start / end)
    out_off = np.empty(len(captions) + 1, dtype=np.int64)
    out_ids = np.empty(cap, dtype=np.int32)
    n = lib.laff_tokenize_lookup(blob, offsets.ctypes.data, len(captions), vocab.handle, stop, int(mode), int(unk_id), int(start_id),
                                 int(end_id), out_off.ctypes.data, out_ids.ctypes.data, cap)
    if n < 0 or n > cap:
        raise LaffError("laff_tokenize_lookup failed (%d): %s" % (n, (lib.laff_last_error() or b"?").decode()))
    return out_off, out_ids[:n]


# ----------------------------------------------------------------------------------------------------------------
# txt2vec
# ----------------------------------------------------------------------------------------------------------------
class Txt2Vec:
    """txt2vec.py:12-47.  norm: 0 none; 1 / 2 are declared by the reference but its `encoding` calls a method that does
    not exist (`self.do_norm`, txt2vec.py:41) — the shipped configs use 0, anything else raises here too."""
    remove_stopword = False

    def __init__(self, data_path, norm=0, clean=True):
        assert norm in [0, 1, 2], "invalid norm %s" % norm
        if norm != 0:
            raise AttributeError("'%s' object has no attribute 'do_norm' (the reference fails the same way for norm > 0)"
                                 % self.__class__.__name__)
        self.data_path = data_path
        self.norm = norm
        self.lang = "en"
        self.clean = clean

    def _preprocess(self, query):
        return TextTool.tokenize(query, clean=self.clean, language=self.lang, remove_stopword=self.remove_stopword)

    def encode_batch(self, captions: Sequence[str], device=None) -> torch.Tensor:
        raise Exception("encoding not implemented yet!")

    def encoding(self, query):
        return self.encode_batch([query])[0].cpu().numpy().astype(np.float64)


class BowVec(Txt2Vec):
    """txt2vec.py:49-86: count of every in-vocabulary token."""

    def __init__(self, data_path, norm=0, clean=True, vocab: Optional[Vocabulary] = None):
        super().__init__(data_path, norm, clean)
        self.vocab = vocab if vocab is not None else load_vocab(data_path)
        self.ndims = len(self.vocab)

    def token_ids(self, query) -> List[int]:
        return [i for i in (self.vocab.find(w) for w in self._preprocess(query)) if i >= 0]

    def token_csr(self, captions):
        """CSR token ids of a batch: the native tokeniser when clean=True (the reference's default), else per caption."""
        if not self.clean:
            lists = [self.token_ids(c) for c in captions]
            off = np.zeros(len(lists) + 1, dtype=np.int64)
            np.cumsum([len(l) for l in lists], out=off[1:])
            return off, np.fromiter((i for l in lists for i in l), dtype=np.int32, count=int(off[-1]))
        if getattr(self, "_native", None) is None:
            words = list(self.vocab.word2idx.keys())
            self._native = NativeVocab(words, [self.vocab.word2idx[w] for w in words])
        return tokenize_lookup(captions, self._native, 1, self.remove_stopword)

    def encode_batch(self, captions, device=None):
        dev = _device(device)
        off, ids = self.token_csr(captions)
        return ops.bow_counts(torch.from_numpy(off).to(dev), torch.from_numpy(ids).to(dev), self.ndims)

    def encode_sparse(self, captions, device=None):
        """The same batch kept sparse: CSR token ids on the device (ops.SparseRows) for the gather-sum projection."""
        dev = _device(device)
        off, ids = self.token_csr(captions)
        return ops.SparseRows(torch.from_numpy(off).to(dev), torch.from_numpy(ids).to(dev), self.ndims)

    def __len__(self):
        return self.ndims


class BowVecNSW(BowVec):
    remove_stopword = True


class W2Vec(Txt2Vec):
    """txt2vec.py:89-112: mean of the word vectors of the distinct in-vocabulary words (`BigFile.read` collapses
    duplicates and orders by file position, bigfile.py:204-211), zeros when no word is known."""

    def __init__(self, data_path, norm=0, clean=True, w2v: Optional[BigFile] = None):
        super().__init__(data_path, norm, clean)
        self.w2v = w2v if w2v is not None else BigFile(data_path)
        _, self.ndims = self.w2v.shape()
        self._table = None

    def table(self, device=None) -> torch.Tensor:
        """The word-vector table resident on the device (uploaded once, streamed from the feature file)."""
        dev = _device(device)
        if self._table is None or self._table.device != dev:
            self._table = self.w2v.to_device(0, self.w2v.nr_of_images, dev)
        return self._table

    def word_ids(self, query) -> List[int]:
        n2i = self.w2v.name2index
        return sorted({n2i[w] for w in self._preprocess(query) if w in n2i})

    def word_csr(self, captions):
        if not self.clean:
            lists = [self.word_ids(c) for c in captions]
            off = np.zeros(len(lists) + 1, dtype=np.int64)
            np.cumsum([len(l) for l in lists], out=off[1:])
            return off, np.fromiter((i for l in lists for i in l), dtype=np.int32, count=int(off[-1]))
        if getattr(self, "_native", None) is None:
            self._native = NativeVocab(self.w2v.names)           # id = row in the vector file
        return tokenize_lookup(captions, self._native, 2, self.remove_stopword)

    def encode_batch(self, captions, device=None):
        dev = _device(device)
        off, ids = self.word_csr(captions)
        return ops.gather_mean(self.table(dev), torch.from_numpy(off).to(dev), torch.from_numpy(ids).to(dev))


class W2VecNSW(W2Vec):
    remove_stopword = True


class IndexVec(Txt2Vec):
    """txt2vec.py:115-128: '<start>' + tokens + '<end>' as vocabulary indices (unknown words -> '<unk>')."""

    def __init__(self, data_path, clean=True, vocab: Optional[Vocabulary] = None):
        super().__init__(data_path, 0, clean)
        self.vocab = vocab if vocab is not None else load_vocab(data_path)
        self.ndims = len(self.vocab)

    def _preprocess(self, query):
        words = TextTool.tokenize(query, clean=self.clean, language=self.lang, remove_stopword=False)
        return ["<start>"] + words + ["<end>"]

    def encoding(self, query):
        return np.array([self.vocab(word) for word in self._preprocess(query)])

    def encoding_batch(self, captions):
        """(ids int32 [B, T] zero padded, lengths int32 [B]) of a batch — what GruTxtEncoder.forward assembles caption by
        caption (model/model.py:345-352), natively."""
        if not self.clean:
            vecs = [self.encoding(c) for c in captions]
            lengths = np.array([len(v) for v in vecs], dtype=np.int32)
            ids = np.zeros((len(vecs), int(lengths.max()) if len(vecs) else 0), dtype=np.int32)
            for i, v in enumerate(vecs):
                ids[i, : lengths[i]] = v
            return ids, lengths
        if getattr(self, "_native", None) is None:
            words = list(self.vocab.word2idx.keys())
            self._native = NativeVocab(words, [self.vocab.word2idx[w] for w in words])
        unk = self.vocab.word2idx.get("<unk>", -1) if "gru" in self.vocab.encoding else -1
        off, flat = tokenize_lookup(captions, self._native, 0, False, unk, self.vocab("<start>"), self.vocab("<end>"))
        if (flat < 0).any():  # textlib.py:107: only a 'gru' vocabulary maps unknown words to <unk>
            raise Exception("word out of vocab")
        lengths = np.diff(off).astype(np.int32)
        ids = np.zeros((len(captions), int(lengths.max()) if len(captions) else 0), dtype=np.int32)
        ids[np.arange(ids.shape[1])[None, :] < lengths[:, None]] = flat
        return ids, lengths


NAME_TO_T2V = {"bow": BowVec, "bow_nsw": BowVecNSW, "w2v": W2Vec, "w2v_nsw": W2VecNSW, "idxvec": IndexVec}


def get_txt2vec(name):
    assert name in NAME_TO_T2V
    return NAME_TO_T2V[name]


# ----------------------------------------------------------------------------------------------------------------
# GRU sentence encoder (numeric part)
# ----------------------------------------------------------------------------------------------------------------
def gru_encode(we: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor, b_ih: torch.Tensor, b_hh: torch.Tensor,
               ids: torch.Tensor, lengths: torch.Tensor, pooling: str = "mean", prepared: Optional[dict] = None):
    """nn.Embedding -> single-layer unidirectional nn.GRU (batch_first, h0 = 0) over packed sequences -> pooling
    (model/model.py:340-387).  ids int32 [B, T] (padding arbitrary beyond lengths[b]), lengths int32 [B].

    The two matrix products run on the tcgen05 GEMM engine with 3-term bf16-split operands (fp32-grade products:
    the recurrence feeds its own rounding back T times, so plain bf16 operands would drift); the input-side
    product is one GEMM over all B*T tokens, the hidden-side one GEMM per step, gates in `laff_gru_cell`."""
    B, T = ids.shape
    H = w_hh.shape[1]
    dev = we.device
    p = prepared if prepared is not None else {}
    if "wih16" not in p:
        p["wih16"] = ops.split3_16(w_ih.detach().float(), 1, torch.bfloat16)
        p["whh16"] = ops.split3_16(w_hh.detach().float(), 1, torch.bfloat16)
        p["b_ih"] = b_ih.detach().float().contiguous()
        p["b_hh"] = b_hh.detach().float().contiguous()
    lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    x = ops.gather_rows(we.detach(), ids.to(dev).reshape(-1))
    gi = ops.project(ops.split3_16(x, 0, torch.bfloat16), p["wih16"], p["b_ih"], "none").view(B, T, 3 * H)
    h = torch.zeros((B, H), dtype=torch.float32, device=dev)
    h2 = torch.empty_like(h)
    want_mean = pooling in ("mean", "mean_last")
    want_last = pooling in ("last", "mean_last")
    if not (want_mean or want_last):
        raise Exception("pooling %s is invalid" % pooling)
    acc = torch.zeros_like(h) if want_mean else None
    last = torch.zeros_like(h) if want_last else None
    gh = torch.empty((B, 3 * H), dtype=torch.float32, device=dev)
    for t in range(T):
        ops.project(ops.split3_16(h, 0, torch.bfloat16), p["whh16"], p["b_hh"], "none", out=gh)
        ops.gru_cell(gi[:, t], gh, h, lengths, t, h2, acc, last)
        h, h2 = h2, h
    if want_mean:
        ops.mean_over_length(acc, lengths)
    if pooling == "mean":
        return acc
    if pooling == "last":
        return last
    return torch.cat((acc, last), dim=1)


# ----------------------------------------------------------------------------------------------------------------
# GRU sentence encoder: training (forward that keeps what BPTT needs, and the backward)
# ----------------------------------------------------------------------------------------------------------------
def _split_ops(x: torch.Tensor, side: int) -> torch.Tensor:
    return ops.split3_16(x, side, torch.bfloat16)


def gru_encode_train(we, w_ih, w_hh, b_ih, b_hh, ids: torch.Tensor, lengths: torch.Tensor, pooling: str = "mean"):
    """gru_encode that keeps the per-step pre-activations and states.  Returns (pooled features, cache)."""
    B, T = ids.shape
    H = w_hh.shape[1]
    dev = we.device
    if pooling not in ("mean", "last", "mean_last"):
        raise Exception("pooling %s is invalid" % pooling)
    lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    ids = ids.to(dev).to(torch.int32).contiguous()
    wih16, whh16 = _split_ops(w_ih.detach().float(), 1), _split_ops(w_hh.detach().float(), 1)
    x = ops.gather_rows(we.detach(), ids.reshape(-1))
    gi = ops.project(_split_ops(x, 0), wih16, b_ih.detach().float().contiguous(), "none").view(B, T, 3 * H)
    h_all = torch.zeros((T + 1, B, H), dtype=torch.float32, device=dev)
    gh_all = torch.empty((T, B, 3 * H), dtype=torch.float32, device=dev)
    acc = torch.zeros((B, H), dtype=torch.float32, device=dev) if pooling != "last" else None
    last = torch.zeros((B, H), dtype=torch.float32, device=dev) if pooling != "mean" else None
    bhh = b_hh.detach().float().contiguous()
    for t in range(T):
        ops.project(_split_ops(h_all[t], 0), whh16, bhh, "none", out=gh_all[t])
        ops.gru_cell(gi[:, t], gh_all[t], h_all[t], lengths, t, h_all[t + 1], acc, last)
    if acc is not None:
        ops.mean_over_length(acc, lengths)
    out = acc if pooling == "mean" else last if pooling == "last" else torch.cat((acc, last), dim=1)
    cache = {"x": x, "ids": ids, "lengths": lengths, "gi": gi, "gh_all": gh_all, "h_all": h_all, "pooling": pooling, "B": B, "T": T,
             "H": H}
    return out, cache


def gru_backward(cache: dict, dout: torch.Tensor, we, w_ih, w_hh, grads: dict) -> None:
    """Backward through time of gru_encode_train.  dout: d loss / d pooled features.  Writes into the preallocated
    gradient tensors grads = {'we', 'w_ih', 'w_hh', 'b_ih', 'b_hh'} (shapes of the parameters; 'we' is zeroed here).
    Per step: one elementwise kernel + one tcgen05 GEMM (dGh_t @ W_hh); afterwards three GEMMs over K = B*T for dW_hh,
    dW_ih and the embedding gradients."""
    B, T, H = cache["B"], cache["T"], cache["H"]
    dev = dout.device
    pooling = cache["pooling"]
    dout = dout.reshape(B, -1).float().contiguous()
    dmean = dout[:, :H].contiguous() if pooling != "last" else None
    dlast = (dout[:, H:] if pooling == "mean_last" else dout).contiguous() if pooling != "mean" else None
    whhT16 = ops.transpose_16(w_hh.detach().float(), torch.bfloat16, 3, 1)       # [H, 3 * 3H]: K-major W_hh^T
    dgi_all = torch.empty((B, T, 3 * H), dtype=torch.float32, device=dev)
    dgh_all = torch.empty((T, B, 3 * H), dtype=torch.float32, device=dev)
    carry = torch.zeros((B, H), dtype=torch.float32, device=dev)
    through = None
    buf = torch.empty((B, H), dtype=torch.float32, device=dev)
    for t in range(T - 1, -1, -1):
        ops.gru_cell_backward(cache["gi"][:, t], cache["gh_all"][t], cache["h_all"][t], dmean, dlast, through, cache["lengths"], t,
                              carry, dgi_all[:, t], dgh_all[t])
        if t > 0:
            through = ops.project(_split_ops(dgh_all[t], 0), whhT16, None, "none", out=buf)
    dgi2 = dgi_all.view(B * T, 3 * H)
    dgh2 = dgh_all.view(T * B, 3 * H)
    hprev2 = cache["h_all"][:T].reshape(T * B, H)
    ops.sim_dense(ops.transpose_16(dgh2, torch.bfloat16, 3, 0), ops.transpose_16(hprev2, torch.bfloat16, 3, 1), 1.0, out=grads["w_hh"])
    ops.column_sum(dgh2, grads["b_hh"])
    ops.sim_dense(ops.transpose_16(dgi2, torch.bfloat16, 3, 0), ops.transpose_16(cache["x"], torch.bfloat16, 3, 1), 1.0, out=grads["w_ih"])
    ops.column_sum(dgi2, grads["b_ih"])
    dx = ops.project(_split_ops(dgi2, 0), ops.transpose_16(w_ih.detach().float(), torch.bfloat16, 3, 1), None, "none")
    grads["we"].zero_()
    ops.scatter_add_rows(dx, cache["ids"].reshape(-1), grads["we"])
