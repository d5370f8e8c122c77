"""Golden vectors for the result-writer / id-based evaluation path, produced by the UNMODIFIED
reference: predictor.txt2video_write_to_file and predictor.write_to_predict_result_file are called as they are; the
two evaluation loops of predictor.get_predict_file (predictor.py:236-246 text->video, :262-270 video->text) are inline
code there, so this script runs the same statements on the same arrays and feeds the reference's evaluation.eval.

    python tests/golden/make_golden_predictor.py      # writes tests/golden/predictor.json (needs /root/reference)

Scores are distinct within every row (no ties) so numpy's unstable default argsort cannot matter, except in the
`tied` case, which is evaluated with kind='stable' (the documented tie rule) and marked as such.
"""
from __future__ import annotations

import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def synth_case(seed, n_vis, caps_per_vis, tied=False):
    rng = np.random.RandomState(seed)
    vis_ids = ["video%03d" % i for i in rng.permutation(n_vis)]
    txt_ids, captions = [], {}
    for v in sorted(vis_ids):
        for c in range(caps_per_vis):
            tid = "%s#enc#%d" % (v, c)
            txt_ids.append(tid)
            captions[tid] = "a caption about %s number %d" % (v, c)
    order = rng.permutation(len(txt_ids))
    txt_ids = [txt_ids[i] for i in order]
    t2i = rng.uniform(-0.3, 0.3, size=(len(txt_ids), n_vis)).astype(np.float32)
    for i, t in enumerate(txt_ids):  # make the ground truth score well, not always best
        t2i[i, vis_ids.index(t.split("#")[0])] += np.float32(rng.uniform(0.0, 0.5))
    if tied:
        t2i = np.round(t2i * 8) / 8  # many exact ties
        t2i = t2i.astype(np.float32)
    return t2i, txt_ids, vis_ids, captions


def reference_eval_loops(reval, t2i, txt_ids, vis_ids, kind=None):
    kw = {} if kind is None else {"kind": kind}
    inds = np.argsort(t2i, axis=1, **kw)
    label_matrix = np.zeros(inds.shape)
    for index in range(inds.shape[0]):  # predictor.py:238-241
        ind = inds[index][::-1]
        gt_index = np.where(np.array(vis_ids)[ind] == txt_ids[index].split('#')[0])[0]
        label_matrix[index][gt_index] = 1
    t2v = reval.eval(label_matrix)
    i2t_matrix = t2i.T  # predictor.py:262-270
    inds = np.argsort(i2t_matrix, axis=1, **kw)
    label_matrix = np.zeros(inds.shape)
    txt_ids2 = [txt_id.split('#')[0] for txt_id in txt_ids]
    for index in range(inds.shape[0]):
        ind = inds[index][::-1]
        label_matrix[index][np.where(np.array(txt_ids2)[ind] == vis_ids[index])[0]] = 1
    v2t = reval.eval(label_matrix)
    return [float(x) for x in t2v], [float(x) for x in v2t]


def main():
    mg.install_shims()
    import predictor as rpred
    import evaluation as reval
    out = {}
    for name, (seed, n_vis, cpv, tied) in {"single_caption": (101, 40, 1, False), "multi_caption": (102, 25, 4, False),
                                           "tied": (103, 30, 3, True)}.items():
        t2i, txt_ids, vis_ids, captions = synth_case(seed, n_vis, cpv, tied)
        case = {"seed": seed, "n_vis": n_vis, "caps_per_vis": cpv, "tied": tied}
        t2v, v2t = reference_eval_loops(reval, t2i, txt_ids, vis_ids, kind="stable" if tied else None)
        case["t2v_metrics"], case["v2t_metrics"] = t2v, v2t
        if not tied:
            inds = np.argsort(t2i, axis=1)
            fake_loader = types.SimpleNamespace(dataset=types.SimpleNamespace(
                get_caption_dict_by_id=lambda tid, c=captions: {"caption": c[tid]}))
            for thr in (16, 2000):  # n_vis >= Threshold (top-16) and n_vis < Threshold (the 0:-1 slice)
                with tempfile.TemporaryDirectory() as td:
                    f = os.path.join(td, "id.sent.score.txt")
                    pk = os.path.join(td, "t2v.pkl")
                    rpred.txt2video_write_to_file(f, inds, vis_ids, txt_ids, t2i, pkl_saved_file=pk, txt_loader=fake_loader,
                                                  Threshold=thr)
                    case["lines_thr%d" % thr] = open(f).read().splitlines()
                    d = pickle.load(open(pk, "rb"))
                    case["pkl_thr%d" % thr] = {k: {"query": v["query"], "rank_list": v["rank_list"],
                                                   "sim_value": [repr(float(x)) for x in v["sim_value"]]} for k, v in d.items()}
        out[name] = case
    # write_to_predict_result_file: everything after the time stamp
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "TextToVideo", "result.txt")
        ck = {"opt": types.SimpleNamespace(parm_adjust_config="0_12_0_12_0_0_1")}
        rpred.write_to_predict_result_file(f, "some/model/path\tcollection", ck, (12.3456, 45.6789, 78.9, 3.0, 25.12345, 0.45678, 0.5),
                                           name_str="Text to video")
        line = open(f).read()
        out["result_file_line_after_timestamp"] = line.split("\t", 1)[1]
    json.dump(out, open(os.path.join(HERE, "predictor.json"), "w"), indent=1)
    print("written", os.path.join(HERE, "predictor.json"))


if __name__ == "__main__":
    main()
