"""Golden vectors for the text front-end from the UNMODIFIED reference: textlib.TextTool.tokenize,
txt2vec.BowVecNSW / W2VecNSW / IndexVec and model.model.GruTxtEncoder / BoWTxtEncoder / W2VTxtEncoder on a small
synthetic vocabulary, word-vector file and caption set.

    python tests/golden/make_golden_text.py    # writes tests/golden/text/{golden.npz, meta.json, vocab_*.pkl, w2v/...}

The stop words the reference applied to these captions are recorded in meta.json (the reference's stopwords_en.txt is
not copied); the vocabulary pickles are written with the reference's own textlib.Vocabulary class so that loading a
reference pickle is exercised by the tests.
"""
import json
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from laff_b200 import synth  # noqa: E402

CAPTIONS = [
    "A man is playing the guitar, while a DOG barks!",
    "two dogs run on the grass",
    "the   the the",                                   # only stop words -> empty bag
    "A cat??? sits on a sofa and a cat sleeps",        # repeated word: counted twice by BoW, once by w2v
    "zebra",                                           # out of every vocabulary
    "someone is cooking pasta in the kitchen with a friend and a dog near the window of the house",
    "Guitar\rplayer 42 plays",
]
WORDS = ("man playing guitar dog barks two dogs run grass cat sits sofa sleeps someone cooking pasta kitchen friend near "
         "window house player 42 plays a is the on and in with of while").split()


def main():
    mg.install_shims()
    import torch
    import textlib
    import txt2vec
    import model.model as mm
    mm.device = torch.device("cpu")
    mm.float16 = False
    out_dir = os.path.join(HERE, "text")
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.RandomState(17)

    # vocabularies written with the reference's class
    bow_vocab = textlib.Vocabulary("bow_nsw")
    for w in WORDS:
        if w not in textlib.ENGLISH_STOP_WORDS:
            bow_vocab.add(w)
    gru_vocab = textlib.Vocabulary("gru")
    for w in ["<pad>", "<start>", "<end>", "<unk>"] + WORDS:
        gru_vocab.add(w)
    pickle.dump(bow_vocab, open(os.path.join(out_dir, "vocab_bow_nsw.pkl"), "wb"))
    pickle.dump(gru_vocab, open(os.path.join(out_dir, "vocab_gru.pkl"), "wb"))
    # word vectors (a BigFile whose names are words, not in vocabulary order)
    w2v_words = [WORDS[i] for i in rng.permutation(len(WORDS))][:28] + ["unused1", "unused2"]
    w2v_dim = 20
    table = rng.standard_normal((len(w2v_words), w2v_dim)).astype(np.float32)
    wd = os.path.join(out_dir, "w2v")
    os.makedirs(wd, exist_ok=True)
    open(os.path.join(wd, "shape.txt"), "w").write("%d %d" % table.shape)
    open(os.path.join(wd, "id.txt"), "w").write("\n".join(w2v_words))
    table.tofile(os.path.join(wd, "feature.bin"))

    t2v_bow = txt2vec.BowVecNSW(os.path.join(out_dir, "vocab_bow_nsw.pkl"))
    t2v_w2v = txt2vec.W2VecNSW(wd)
    t2v_idx = txt2vec.IndexVec(os.path.join(out_dir, "vocab_gru.pkl"))
    meta = {"captions": CAPTIONS, "tokens": [], "tokens_nsw": [], "index": []}
    stop_used = set()
    for c in CAPTIONS:
        a = textlib.TextTool.tokenize(c, clean=True, language="en")
        b = textlib.TextTool.tokenize(c, clean=True, language="en", remove_stopword=True)
        meta["tokens"].append(a)
        meta["tokens_nsw"].append(b)
        stop_used |= set(a) - set(b)
        meta["index"].append([int(x) for x in t2v_idx.encoding(c)])
    meta["stopwords_used"] = sorted(stop_used)
    meta["bow_words"] = [bow_vocab[i] for i in range(len(bow_vocab))]
    meta["gru_words"] = [gru_vocab[i] for i in range(len(gru_vocab))]
    out = {"w2v_table": table}
    out["bow_enc"] = np.stack([t2v_bow.encoding(c) for c in CAPTIONS])
    out["w2v_enc"] = np.stack([t2v_w2v.encoding(c) for c in CAPTIONS])

    # the nn.Module encoders
    opt = types.SimpleNamespace(t2v_bow=t2v_bow, t2v_w2v=t2v_w2v, t2v_idx=t2v_idx, rnn_layer=1)
    with torch.no_grad():
        out["bow_module"] = mm.BoWTxtEncoder(opt)({"caption": CAPTIONS})["text_features"].numpy()
        out["w2v_module"] = mm.W2VTxtEncoder(opt)({"caption": CAPTIONS})["text_features"].numpy()
        for tag, we_dim, H in (("small", 12, 32), ("full", 500, 1024)):
            for pooling in ("mean", "last", "mean_last"):
                if tag == "full" and pooling != "mean":
                    continue
                opt.we_dim, opt.rnn_size, opt.pooling = we_dim, H, pooling
                opt.we = torch.zeros(len(gru_vocab), we_dim)
                enc = mm.GruTxtEncoder(opt).eval()
                sd = {k: torch.from_numpy(np.asarray(synth.param(71, "gru_%s/%s" % (tag, k), tuple(v.shape))))
                      for k, v in enc.state_dict().items()}
                enc.load_state_dict(sd)
                if tag == "small":
                    for k, v in sd.items():
                        out["gru_small/%s" % k] = v.numpy()
                out["gru_%s_%s" % (tag, pooling)] = enc({"caption": CAPTIONS})["text_features"].numpy()
    np.savez_compressed(os.path.join(out_dir, "golden.npz"), **out)
    json.dump(meta, open(os.path.join(out_dir, "meta.json"), "w"), indent=1)
    print("written", out_dir)


if __name__ == "__main__":
    main()
