"""Imports the staged, unmodified reference (baseline/_ref/laff_reference, see stage_reference.py) on a box that has
neither the reference tree nor its optional dependencies -- TEST / BASELINE INFRASTRUCTURE (bench.py's CPU arm).

Leaf imports that are absent are shimmed exactly as tests/golden/make_golden.py does (ftfy, nltk, prefetch_generator,
torchvision's removed Kinetics400 alias); the four text encoders are replaced by pass-through modules because
CLIPEncoder.__init__ downloads weights (model/model.py:483) and the text features are inputs at this tier.  Everything
the arm times -- get_txt2vis_matrix / compute_sim / cosine_sim / l2norm, np.argsort, evaluation.eval -- is the
reference's own code.
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "laff_reference")


def available() -> bool:
    return os.path.exists(os.path.join(STAGED, "MANIFEST.json"))


def _shims():
    os.environ.setdefault("HOME", "/tmp")
    for name in ("ftfy", "prefetch_generator"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.fix_text = lambda s: s
                m.BackgroundGenerator = object
                sys.modules[name] = m
    try:
        import nltk  # noqa: F401
    except Exception:
        nltk = types.ModuleType("nltk")
        nltk.word_tokenize = lambda s: s.split()
        nltk.pos_tag = lambda toks: [(t, "NN") for t in toks]
        stem = types.ModuleType("nltk.stem")
        stem.WordNetLemmatizer = object
        corpus = types.ModuleType("nltk.corpus")
        corpus.stopwords = types.SimpleNamespace(words=lambda lang: [])
        corpus.wordnet = types.SimpleNamespace()
        nltk.stem, nltk.corpus = stem, corpus
        sys.modules.update({"nltk": nltk, "nltk.stem": stem, "nltk.corpus": corpus})
    try:
        import torchvision.datasets as tvd
        if not hasattr(tvd, "Kinetics400"):
            tvd.Kinetics400 = getattr(tvd, "Kinetics", object)
    except Exception:
        pass


def load():
    """-> (model.model module, evaluation module) of the staged reference, device = cpu, float16 off."""
    if not available():
        raise ImportError("no staged reference under %s (run oracle/stage_reference.py where /root/reference exists)" % STAGED)
    _shims()
    if STAGED not in sys.path:
        sys.path.insert(0, STAGED)
    import torch
    import model.model as mm  # noqa  (the staged reference)
    import evaluation as reval
    mm.device = torch.device("cpu")
    mm.float16 = False

    class PassThrough(torch.nn.Module):
        key = None

        def __init__(self, opt=None):
            super().__init__()

        def forward(self, caption_feat_dict, task3=False):
            return {"text_features": caption_feat_dict[self.key]}

    for cls, key in (("GruTxtEncoder", "gru"), ("BoWTxtEncoder", "bow"), ("W2VTxtEncoder", "w2v"), ("CLIPEncoder", "clip")):
        setattr(mm, cls, type("PassThrough_" + key, (PassThrough,), {"key": key}))
    return mm, reval
