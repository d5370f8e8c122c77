"""End to end on an on-disk synthetic collection in the reference's layout (SURVEY §8f N1 + N2 + N3):
feature files -> resident gallery index, caption strings + precomputed CLIP features -> queries, ranking, metric lines,
t2v.pkl and id.sent.score.txt — laff_b200.collection.predict_collection, the body of predictor.get_predict_file."""
import json
import os
import pickle
import types

import numpy as np
import pytest
import torch

from helpers import load_numpy_state
from laff_b200 import config as cfg
from laff_b200 import model as M
from laff_b200 import predictor as P
from laff_b200 import synth
from laff_b200 import text as T
from laff_b200.bigfile import write_bigfile
from laff_b200.collection import predict_collection, read_captions
from oracle import laff_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TXT = os.path.join(HERE, "golden", "text")
META = json.load(open(os.path.join(TXT, "meta.json")))


def make_collection(root, coll, V, caps_per_vid, dims, clip_dim, rng):
    base = os.path.join(root, coll)
    vis_ids = ["video%04d" % i for i in rng.permutation(V)]
    os.makedirs(os.path.join(base, "VideoSets"))
    open(os.path.join(base, "VideoSets", coll + ".txt"), "w").write("\n".join(vis_ids) + "\n")
    feats = {}
    for name, d in dims.items():
        x = rng.standard_normal((V, d)).astype(np.float32)
        feats[name] = x
        order = rng.permutation(V)
        write_bigfile(os.path.join(base, "FeatureData", name), [vis_ids[j] for j in order], x[order])
    words = META["bow_words"]
    lines, cap_ids = [], []
    for v in sorted(vis_ids):
        for c in range(caps_per_vid):
            cid = "%s#enc#%d" % (v, c)
            cap_ids.append(cid)
            lines.append("%s A %s and the %s, %s!" % (cid, *rng.choice(words, 3)))
    os.makedirs(os.path.join(base, "TextData"))
    open(os.path.join(base, "TextData", coll + ".caption.txt"), "w").write("\n".join(lines) + "\n\n")
    open(os.path.join(base, "TextData", "simple_query.txt"), "w").write("q1 a dog runs\nq2\nq3 two cats on the sofa\n")
    clip = rng.standard_normal((len(cap_ids) + 3, clip_dim)).astype(np.float32)
    write_bigfile(os.path.join(base, "TextData", "CLIP_feats"), cap_ids + ["q1", "q2", "q3"], clip)
    return vis_ids, feats, cap_ids, dict(zip(cap_ids + ["q1", "q2", "q3"], clip))


def test_predict_collection_end_to_end(tmp_path):
    T.TextTool.set_stopwords(META["stopwords_used"])
    try:
        rng = np.random.RandomState(4)
        bow = T.BowVecNSW(os.path.join(TXT, "vocab_bow_nsw.pkl"))
        w2v = T.W2VecNSW(os.path.join(TXT, "w2v"))
        idx = T.IndexVec(os.path.join(TXT, "vocab_gru.pkl"))
        dims = dict(synth.DIMS)
        dims.update(bow=bow.ndims, w2v=w2v.ndims)
        c = cfg.laff_config(4096, 8, dims)
        c.t2v_bow, c.t2v_w2v, c.t2v_idx = bow, w2v, idx
        c.we_dim, c.rnn_size, c.rnn_layer, c.we = 500, 1024, 1, None
        c.text_encoding["CLIP_encoding"]["dir_name"] = "CLIP_feats"
        # base_config.py:171-173 ships a non-empty vid_frame_feats with frame_feat_input = False: plain LAFF configs (and the
        # config stored in a reference checkpoint) carry both, and only the flag may decide (predictor.py:191)
        c.vid_frame_feats, c.frame_feat_input = [synth.VIS_FRAME], False
        model = M.get_model("LAFF", torch.device("cuda"), c)
        load_numpy_state(model, {k: np.asarray(synth.param(5, k, tuple(v.shape))) for k, v in model.state_dict().items()})
        V, cpv = 60, 2
        vis_ids, feats, cap_ids, clip = make_collection(str(tmp_path), "toyset", V, cpv, dict(c.vis_fc_layers[0]), 512, rng)
        ck = {"opt": types.SimpleNamespace(parm_adjust_config="0_12_0_12_0_0_1")}
        prf = str(tmp_path / "results" / "pred.txt")
        res = predict_collection(model, c, str(tmp_path), "toyset", ["toyset.caption.txt", "simple_query.txt"], "laff_sim",
                                 predict_result_file=prf, model_path="ckpt.pth", checkpoint=ck)
        # --- the same scores computed piecewise through the public model API
        ids2, caps = read_captions(str(tmp_path / "toyset" / "TextData" / "toyset.caption.txt"))
        assert ids2 == cap_ids
        from laff_b200 import loss as L
        from laff_b200 import ops
        with torch.no_grad():
            dt = L.operand_dtype()
            v, v16 = model.vis_net.encode({k: torch.from_numpy(x) for k, x in feats.items()}, out16_dtype=dt)
            t, t16 = model.txt_net.encode({"caption": [caps[i] for i in cap_ids],
                                           "CLIP_encoding": torch.from_numpy(np.stack([clip[i] for i in cap_ids]))}, out16_dtype=dt)
            # the collection path ranks the 16-bit embeddings the fused kernel writes; get_txt2vis_matrix re-normalises the
            # fp32 copies before rounding, which can move a component by one 16-bit ulp: same scores within that rounding
            s = ops.sim_dense(t16.reshape(len(cap_ids), -1), v16.reshape(V, -1), 1.0 / 8).cpu().numpy()
            assert np.abs(model.get_txt2vis_matrix(t, v).cpu().numpy() - s).max() <= 2e-5
        t2v, _ = O.predictor_t2v_eval(s, cap_ids, vis_ids)
        v2t, _ = O.predictor_v2t_eval(s, cap_ids, vis_ids)
        np.testing.assert_allclose(res["toyset.caption.txt"]["t2v"], t2v, atol=1e-9)
        np.testing.assert_allclose(res["toyset.caption.txt"]["v2t"], v2t, atol=1e-9)
        out_dir = tmp_path / "toyset" / "SimilarityIndex" / "toyset.caption.txt" / "laff_sim"
        d = pickle.load(open(out_dir / "t2v.pkl", "rb"))
        ref = O.t2v_shot_dict(s, cap_ids, vis_ids, caps, 500)
        assert list(d.keys()) == cap_ids and all(d[k]["rank_list"] == ref[k]["rank_list"] and d[k]["query"] == caps[k] for k in d)
        for sub in ("TextToVideo", "VideoToText"):
            line = open(tmp_path / "results" / sub / "pred.txt").read()
            assert "ckpt.pth\ttoyset\t" in line and line.rstrip("\n").endswith("0\t12\t0\t12\t0\t0\t1")
        # --- ad-hoc queries: id.sent.score.txt (top-2000 rule => V - 1 entries here), empty caption tolerated
        f = tmp_path / "toyset" / "SimilarityIndex" / "simple_query.txt" / "laff_sim" / "id.sent.score.txt"
        lines = f.read_text().splitlines()
        assert [l.split()[0] for l in lines] == ["q1", "q2", "q3"] and all(len(l.split()) == 1 + 2 * (V - 1) for l in lines)
        assert set(lines[0].split()[1::2]) <= set(vis_ids)
        # --- and the no-dense-matrix path gives the same files and text->video metrics
        res2 = predict_collection(model, c, str(tmp_path), "toyset", ["toyset.caption.txt"], "laff_sim_big", predict_result_file=prf,
                                  model_path="ckpt.pth", checkpoint=ck, dense_limit=0)
        np.testing.assert_allclose(res2["toyset.caption.txt"]["t2v"], t2v, atol=1e-9)
        d2 = pickle.load(open(tmp_path / "toyset" / "SimilarityIndex" / "toyset.caption.txt" / "laff_sim_big" / "t2v.pkl", "rb"))
        assert all(d2[k]["rank_list"] == d[k]["rank_list"] for k in d)
    finally:
        T.TextTool._stopwords = None
