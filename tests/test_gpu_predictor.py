"""GPU parity of the result-writer path: laff_topk_dense / laff_rank_multi_gt / laff_multi_gt_metrics and
laff_b200.predictor against the oracle and the reference's own outputs (tests/golden/predictor.json)."""
import json
import os
import pickle
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_predictor import synth_case  # noqa: E402

from laff_b200 import ops  # noqa: E402
from laff_b200 import predictor as P  # noqa: E402
from oracle import laff_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(HERE, "golden", "predictor.json")))
CASES = [k for k in GOLD if isinstance(GOLD[k], dict)]


def case_inputs(name):
    c = GOLD[name]
    return synth_case(c["seed"], c["n_vis"], c["caps_per_vis"], c["tied"])


def scores_with_ties(rng, rows, cols, levels):
    """Rows with heavy exact ties (quantised) plus a few distinct extremes, negative values and zeros of both signs."""
    s = rng.standard_normal((rows, cols)).astype(np.float32)
    if levels:
        s = (np.round(s * levels) / levels).astype(np.float32)
    s[0, : min(cols, 7)] = [0.0, -0.0, 1e-30, -1e-30, 3.0, -3.0, 0.0][: min(cols, 7)]
    return s


@pytest.mark.parametrize("rows,cols,k,levels", [
    (5, 1, 1, 0), (7, 33, 33, 4), (9, 1000, 999, 0), (6, 2990, 500, 8), (3, 16384, 2048, 16),   # whole-row sort path
    (4, 16385, 2000, 0), (5, 50000, 2048, 6), (3, 200000, 500, 3), (2, 1000003, 2000, 2),       # radix-select path
    (3, 40000, 7, 1),
])
def test_topk_dense_is_stable_argsort_reversed(rows, cols, k, levels):
    rng = np.random.RandomState(rows * 131 + cols)
    s = scores_with_ties(rng, rows, cols, levels)
    tv, ti = ops.topk_dense(torch.from_numpy(s).cuda(), k)
    ref_v, ref_i = O.tie_rule_topk(s, k)
    assert np.array_equal(ti.cpu().numpy().astype(np.int64), ref_i)          # integer output: bit-exact
    assert np.array_equal(tv.cpu().numpy(), ref_v)                            # values identical (0.0 == -0.0)


def test_topk_dense_more_than_candidates_and_scale():
    s = np.array([[0.5, -1.0, 0.5], [2.0, 2.0, 2.0]], dtype=np.float32)
    tv, ti = ops.topk_dense(torch.from_numpy(s).cuda(), 5, scale=0.5)
    assert ti.cpu().tolist() == [[2, 0, 1, -1, -1], [2, 1, 0, -1, -1]]
    assert tv.cpu().tolist()[0][:3] == [0.25, 0.25, -0.5] and all(np.isneginf(tv.cpu().numpy()[:, 3:]).ravel())
    with pytest.raises(ops.LaffError):
        ops.topk_dense(torch.zeros(2, 5000, device="cuda"), 4096)


def test_topk_dense_infinities_and_empty_inputs():
    s = np.array([[0.0, np.inf, -np.inf, 1.0, np.inf, -np.inf]], dtype=np.float32)
    for cols_pad in (0, 20000):                      # whole-row sort path and radix-select path
        row = np.concatenate([s, np.full((1, cols_pad), -2.0, np.float32)], 1)
        tv, ti = ops.topk_dense(torch.from_numpy(row).cuda(), 6 if cols_pad == 0 else 5)
        ref_v, ref_i = O.tie_rule_topk(row, tv.shape[1])
        assert np.array_equal(ti.cpu().numpy().astype(np.int64), ref_i) and np.array_equal(tv.cpu().numpy(), ref_v)
    tv, ti = ops.topk_dense(torch.zeros(3, 0, device="cuda"), 4)
    assert tv.shape == (3, 4) and bool(torch.isinf(tv).all()) and bool((ti == -1).all())
    tv, ti = ops.topk_dense(torch.zeros(0, 7, device="cuda"), 4)
    assert tv.shape == (0, 4)


def test_topk_dense_merges_shard_lists_by_global_index():
    """Per-shard lists concatenated (as all_gather delivers them) -> the global list, ties resolved by global index."""
    rng = np.random.RandomState(5)
    Q, V, W, k = 6, 900, 3, 200
    s = scores_with_ties(rng, Q, V, 5)
    ref_v, ref_i = O.tie_rule_topk(s, k)
    bounds = [0, 250, 650, 900]
    vals, idxs = [], []
    for w in range(W):
        lo, hi = bounds[w], bounds[w + 1]
        v, i = O.tie_rule_topk(s[:, lo:hi], k)
        pad = k - v.shape[1]
        vals.append(np.pad(v, ((0, 0), (0, pad)), constant_values=-np.inf))
        idxs.append(np.pad(i + lo, ((0, 0), (0, pad)), constant_values=-1))
    cat_v = torch.from_numpy(np.concatenate(vals, 1)).cuda()
    cat_i = torch.from_numpy(np.concatenate(idxs, 1).astype(np.int32)).cuda()
    tv, ti = ops.topk_dense(cat_v, k, idx_in=cat_i)
    assert np.array_equal(ti.cpu().numpy().astype(np.int64), ref_i)
    assert np.array_equal(tv.cpu().numpy(), ref_v)


@pytest.mark.parametrize("name", CASES)
def test_evaluation_both_directions_match_reference(name):
    t2i, txt_ids, vis_ids, _ = case_inputs(name)
    t2v, rank0 = P.evaluate_t2v(torch.from_numpy(t2i).cuda(), txt_ids, vis_ids)
    v2t, first, ap = P.evaluate_v2t(torch.from_numpy(t2i).cuda(), txt_ids, vis_ids)
    np.testing.assert_allclose(t2v, GOLD[name]["t2v_metrics"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(v2t, GOLD[name]["v2t_metrics"], rtol=0, atol=1e-12)
    _, label = O.predictor_t2v_eval(t2i, txt_ids, vis_ids)
    assert np.array_equal(rank0.cpu().numpy(), label.argmax(1))             # integer ranks: bit-exact
    _, label = O.predictor_v2t_eval(t2i, txt_ids, vis_ids)
    assert np.array_equal(first.cpu().numpy(), label.argmax(1))


def test_rank_multi_gt_ragged_lists_and_out_of_range():
    rng = np.random.RandomState(9)
    R, Cn = 37, 5000
    s = scores_with_ties(rng, R, Cn, 4)
    lists = [sorted(rng.choice(Cn, size=rng.randint(1, 40), replace=False).tolist()) for _ in range(R)]
    off = np.zeros(R + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(l) for l in lists])
    cols = np.concatenate(lists).astype(np.int32)
    ranks = ops.rank_multi_gt(torch.from_numpy(s).cuda(), torch.from_numpy(off).cuda(), torch.from_numpy(cols).cuda()).cpu().numpy()
    order = O.sorted_desc(s)
    pos = np.empty_like(order)
    for r in range(R):
        pos[r, order[r]] = np.arange(Cn)
    ref = np.concatenate([pos[r, lists[r]] for r in range(R)])
    assert np.array_equal(ranks, ref)
    label = np.zeros((R, Cn))
    for r in range(R):
        label[r, pos[r, lists[r]]] = 1
    m, first, ap = ops.multi_gt_metrics(torch.from_numpy(ranks).cuda(), torch.from_numpy(off).cuda())
    np.testing.assert_allclose(m.cpu().numpy()[:7], O.eval_label_matrix(label), rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", [c for c in CASES if not GOLD[c]["tied"]])
@pytest.mark.parametrize("thr", [16, 2000])
def test_writer_files_match_reference(name, thr, tmp_path):
    t2i, txt_ids, vis_ids, captions = case_inputs(name)
    f, pk = tmp_path / "id.sent.score.txt", tmp_path / "t2v.pkl"
    P.txt2video_write_to_file(str(f), None, vis_ids, txt_ids, torch.from_numpy(t2i).cuda(), pkl_saved_file=str(pk),
                              Threshold=thr, captions=captions)
    assert f.read_text().splitlines() == GOLD[name]["lines_thr%d" % thr]
    d = pickle.load(open(pk, "rb"))
    g = GOLD[name]["pkl_thr%d" % thr]
    assert list(d.keys()) == list(g.keys())
    for k in d:
        assert d[k]["query"] == g[k]["query"] and d[k]["rank_list"] == g[k]["rank_list"]
        assert [repr(float(x)) for x in d[k]["sim_value"]] == g[k]["sim_value"]


def test_evaluate_and_write_end_to_end(tmp_path):
    import types
    t2i, txt_ids, vis_ids, captions = case_inputs("multi_caption")
    ck = {"opt": types.SimpleNamespace(parm_adjust_config="0_12_0_12_0_0_1")}
    out = P.evaluate_and_write(torch.from_numpy(t2i).cuda(), txt_ids, vis_ids, str(tmp_path / "out"),
                               str(tmp_path / "results" / "pred.txt"), "model\tcoll", ck, captions=captions)
    np.testing.assert_allclose(out["t2v"], GOLD["multi_caption"]["t2v_metrics"], atol=1e-12)
    np.testing.assert_allclose(out["v2t"], GOLD["multi_caption"]["v2t_metrics"], atol=1e-12)
    assert (tmp_path / "results" / "TextToVideo" / "pred.txt").exists()
    assert (tmp_path / "results" / "VideoToText" / "pred.txt").exists()
    d = pickle.load(open(tmp_path / "out" / "t2v.pkl", "rb"))
    assert len(d) == len(txt_ids) and len(next(iter(d.values()))["rank_list"]) == len(vis_ids) - 1
    out = P.evaluate_and_write(torch.from_numpy(t2i).cuda(), txt_ids, vis_ids, str(tmp_path / "adhoc"), None, "m", ck,
                               captions=captions, with_ground_truth=False)
    lines = open(out["pred_result_file"]).read().splitlines()
    assert lines == O.txt2video_lines(t2i, txt_ids, vis_ids, 2000)
