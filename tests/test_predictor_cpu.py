"""CPU tests for the result-writer / id-based evaluation path: the oracle restatement of
predictor.py:53-88, :236-270 against outputs of the unmodified reference (tests/golden/predictor.json, written by
tests/golden/make_golden_predictor.py), and the host-side id logic of laff_b200.predictor."""
import json
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_predictor import synth_case  # noqa: E402  (pure numpy; the reference is only needed to regenerate)

from laff_b200 import predictor as P  # noqa: E402
from oracle import laff_oracle as O  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "predictor.json")))
CASES = [k for k in GOLD if isinstance(GOLD[k], dict)]


def case_inputs(name):
    c = GOLD[name]
    return synth_case(c["seed"], c["n_vis"], c["caps_per_vis"], c["tied"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_eval_loops_match_reference(name):
    t2i, txt_ids, vis_ids, _ = case_inputs(name)
    t2v, _ = O.predictor_t2v_eval(t2i, txt_ids, vis_ids)
    v2t, _ = O.predictor_v2t_eval(t2i, txt_ids, vis_ids)
    np.testing.assert_allclose(t2v, GOLD[name]["t2v_metrics"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(v2t, GOLD[name]["v2t_metrics"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", [c for c in CASES if not GOLD[c]["tied"]])
@pytest.mark.parametrize("thr", [16, 2000])
def test_oracle_writer_lines_and_pkl_match_reference(name, thr):
    t2i, txt_ids, vis_ids, captions = case_inputs(name)
    assert O.txt2video_lines(t2i, txt_ids, vis_ids, thr) == GOLD[name]["lines_thr%d" % thr]
    d = O.t2v_shot_dict(t2i, txt_ids, vis_ids, captions, thr)
    g = GOLD[name]["pkl_thr%d" % thr]
    assert list(d.keys()) == list(g.keys())
    for k in d:
        assert d[k]["query"] == g[k]["query"] and d[k]["rank_list"] == g[k]["rank_list"]
        assert [repr(float(x)) for x in d[k]["sim_value"]] == g[k]["sim_value"]
    # the 0:-1 slice below the threshold: one video fewer than the gallery holds
    n = len(GOLD[name]["lines_thr%d" % thr][0].split()) // 2
    assert n == (thr if len(vis_ids) >= thr else len(vis_ids) - 1) == P.writer_topk(len(vis_ids), thr)


def test_gt_index_and_caption_lists():
    t2i, txt_ids, vis_ids, _ = case_inputs("multi_caption")
    gt = P.gt_index(txt_ids, vis_ids)
    assert gt.dtype == np.int32 and all(vis_ids[g] == t.split("#")[0] for g, t in zip(gt, txt_ids))
    off, cols = P.caption_lists(txt_ids, vis_ids)
    assert off[0] == 0 and off[-1] == len(txt_ids) and np.all(np.diff(off) == 4)
    for v in range(len(vis_ids)):
        mine = cols[off[v]:off[v + 1]]
        assert list(mine) == sorted(mine) and all(txt_ids[c].split("#")[0] == vis_ids[v] for c in mine)
    with pytest.raises(IndexError):
        P.gt_index(["nosuchvideo#0"], vis_ids)
    with pytest.raises(P.LaffError):
        P.gt_index(txt_ids, vis_ids + [vis_ids[0]])


def test_result_file_line_format(tmp_path):
    f = tmp_path / "TextToVideo" / "result.txt"
    ck = {"opt": types.SimpleNamespace(parm_adjust_config="0_12_0_12_0_0_1")}
    P.write_to_predict_result_file(str(f), "some/model/path\tcollection", ck,
                                   (12.3456, 45.6789, 78.9, 3.0, 25.12345, 0.45678, 0.5), name_str="Text to video")
    assert f.read_text().split("\t", 1)[1] == GOLD["result_file_line_after_timestamp"]
