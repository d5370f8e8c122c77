"""Evaluation / result-writer path of the reference's predictor.py and trainer.validate on the device.

The reference (predictor.py:232-284, trainer.py:579-607) argsorts the whole Q x V matrix on the host, then for every
query materialises `np.array(vis_ids)[ind]` and string-compares it with the caption's video id.  Here the ids are
resolved ONCE into integer ground-truth columns, ranks come from compare-and-count kernels (laff_rank_from_scores /
laff_rank_multi_gt), metrics from laff_rank_metrics / laff_multi_gt_metrics, and the written lists from laff_topk_dense
(top-500 for t2v.pkl, top-2000 for id.sent.score.txt) — nothing of size Q x V ever reaches the host.

Function names, argument order, file formats and quirks follow the reference:
  txt2video_write_to_file      predictor.py:53-88   (TopK = Threshold if len(vis_ids) >= Threshold else `0:-1`)
  write_to_predict_result_file predictor.py:91-126
  evaluate_t2v / evaluate_v2t  predictor.py:236-246 / :262-270 (inline loops there)
  validate                     trainer.py:579-607
Tie order = the documented rule (score desc, index desc), see DESIGN.md §5.
"""
from __future__ import annotations

import os
import pickle
import time
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from ._capi import LaffError


# ----------------------------------------------------------------------------------------------------------------
# ids -> integer ground truth (built once per collection instead of once per query)
# ----------------------------------------------------------------------------------------------------------------
def _vis_position(vis_ids: Sequence[str]) -> Dict[str, int]:
    pos: Dict[str, int] = {}
    for j, v in enumerate(vis_ids):
        if v in pos:
            raise LaffError("duplicate video id %r in vis_ids: the id-to-column map needs unique gallery ids" % (v,))
        pos[v] = j
    return pos


def gt_index(txt_ids: Sequence[str], vis_ids: Sequence[str]) -> np.ndarray:
    """int32 [Q]: column of the video named by txt_id.split('#')[0] (predictor.py:240)."""
    pos = _vis_position(vis_ids)
    out = np.empty(len(txt_ids), dtype=np.int32)
    for i, t in enumerate(txt_ids):
        v = t.split("#")[0]
        if v not in pos:  # the reference fails with IndexError on gt_index[0]
            raise IndexError("caption %r: video %r is not in the gallery" % (t, v))
        out[i] = pos[v]
    return out


def caption_lists(txt_ids: Sequence[str], vis_ids: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """CSR lists of the captions (columns of t2i.T) that belong to every video (predictor.py:265-269):
    (offsets int64 [V + 1], cols int32 [nnz])."""
    pos = _vis_position(vis_ids)
    owner = np.full(len(txt_ids), -1, dtype=np.int64)
    for i, t in enumerate(txt_ids):
        owner[i] = pos.get(t.split("#")[0], -1)
    keep = np.nonzero(owner >= 0)[0]
    order = keep[np.argsort(owner[keep], kind="stable")]
    counts = np.bincount(owner[keep], minlength=len(vis_ids))
    offsets = np.zeros(len(vis_ids) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    return offsets, order.astype(np.int32)


def _metrics_tuple(m: torch.Tensor) -> Tuple[float, ...]:
    m = m.cpu().tolist()
    return (m[0], m[1], m[2], m[3], m[4], m[5], m[6])  # r1, r5, r10, medr, meanr, mir, mAP


def _as_scores(t2i_matrix) -> torch.Tensor:
    t = torch.as_tensor(t2i_matrix)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise LaffError("laff_b200.predictor needs a CUDA device (no CPU fallback)")
        t = t.cuda()
    return t.float()


# ----------------------------------------------------------------------------------------------------------------
# evaluation in both directions
# ----------------------------------------------------------------------------------------------------------------
def evaluate_t2v(t2i_matrix, txt_ids: Sequence[str], vis_ids: Sequence[str]):
    """Text -> video (predictor.py:236-246): returns ((r1, r5, r10, medr, meanr, mir, mAP), rank0 int32 [Q] on device)."""
    s = _as_scores(t2i_matrix)
    gt = torch.from_numpy(gt_index(txt_ids, vis_ids)).to(s.device)
    rank0, _, _ = ops.rank_from_scores(s, gt, 0)
    return _metrics_tuple(ops.rank_metrics(rank0)), rank0


def evaluate_v2t(t2i_matrix, txt_ids: Sequence[str], vis_ids: Sequence[str]):
    """Video -> text (predictor.py:262-270): rows = videos, ground truths = all captions of the video.  Returns
    ((r1, ..., mAP), first int32 [V], ap float64 [V])."""
    s = _as_scores(t2i_matrix)
    offsets, cols = caption_lists(txt_ids, vis_ids)
    if np.any(np.diff(offsets) == 0):
        v = int(np.nonzero(np.diff(offsets) == 0)[0][0])
        raise IndexError("video %r has no caption in txt_ids" % (vis_ids[v],))  # reference: rank[0] on an empty array
    i2t = s.t().contiguous()
    off_d, col_d = torch.from_numpy(offsets).to(s.device), torch.from_numpy(cols).to(s.device)
    ranks = ops.rank_multi_gt(i2t, off_d, col_d)
    m, first, ap = ops.multi_gt_metrics(ranks, off_d)
    return _metrics_tuple(m), first, ap


# ----------------------------------------------------------------------------------------------------------------
# writers
# ----------------------------------------------------------------------------------------------------------------
def writer_topk(n_vis: int, Threshold: int) -> int:
    """Entries per query the reference writes: Threshold when the gallery has at least that many videos, else the slice
    `[0:-1]` = all but the lowest-ranked video (predictor.py:55-58, :64)."""
    return Threshold if n_vis >= Threshold else max(n_vis - 1, 0)


def ranked_lists(t2i_matrix, k: int, query_chunk: int = 4096):
    """(values float32 [Q, k], indices int32 [Q, k]) on the host, best first, by laff_topk_dense."""
    s = _as_scores(t2i_matrix)
    vals, idxs = [], []
    for lo in range(0, s.shape[0], query_chunk):
        v, i = ops.topk_dense(s[lo:lo + query_chunk], k)
        vals.append(v.cpu())
        idxs.append(i.cpu())
    if not vals:
        return np.zeros((0, k), np.float32), np.zeros((0, k), np.int32)
    return torch.cat(vals).numpy(), torch.cat(idxs).numpy()


def txt2video_write_to_file(pred_result_file, inds, vis_ids, txt_ids, t2i_matrix, pkl_saved_file=None, txt_loader=None,
                            Threshold=2000, captions: Optional[Mapping[str, str]] = None):
    """predictor.py:53-88.  `inds` is accepted for signature compatibility: None (lists are extracted on the device
    from `t2i_matrix`) or an already extracted (values, indices) pair, best first, at least writer_topk() wide.
    Writes '<txt_id> <vis_id> <score> <vis_id> <score> ...' lines and/or the t2v.pkl dict
    {txt_id: {'query', 'rank_list', 'sim_value'}}."""
    start = time.time()
    k = writer_topk(len(vis_ids), Threshold)
    if k > ops.MAX_TOPK_DENSE:
        raise LaffError("txt2video_write_to_file: %d entries per query exceed the device list width %d" % (k, ops.MAX_TOPK_DENSE))
    if isinstance(inds, tuple):
        vals, idx = np.asarray(inds[0], dtype=np.float32)[:, :k], np.asarray(inds[1])[:, :k]
    elif k > 0:
        vals, idx = ranked_lists(t2i_matrix, k)
    else:
        vals = np.zeros((len(txt_ids), 0), np.float32)
        idx = np.zeros((len(txt_ids), 0), np.int32)

    def caption_of(tid):
        if captions is not None:
            return captions[tid]
        return txt_loader.dataset.get_caption_dict_by_id(tid)["caption"]

    shot_dict = {}
    if pred_result_file is not None:
        with open(pred_result_file, "w") as fout:
            for q in range(len(txt_ids)):
                fout.write(txt_ids[q] + " " + " ".join([vis_ids[j] + " %s" % v for j, v in zip(idx[q], vals[q])]) + "\n")
    if pkl_saved_file is not None:
        for q in range(len(txt_ids)):
            shot_dict[txt_ids[q]] = {"query": caption_of(txt_ids[q]), "rank_list": [vis_ids[j] for j in idx[q]],
                                     "sim_value": [v for v in vals[q]]}
        with open(pkl_saved_file, "wb") as f:
            pickle.dump(shot_dict, f)
    print("writing result into file time: %.3f seconds\n" % (time.time() - start))
    print("Save to ", pkl_saved_file)


def write_to_predict_result_file(predict_result_file, model_path, checkpoint, result_tuple, name_str="Text to video"):
    """predictor.py:91-126: appends one tab-separated line (time, model path, rounded metrics, parm_adjust_config)."""
    result_file_dir = os.path.dirname(predict_result_file)
    print("pkl result_file_dir: ", predict_result_file)
    if not os.path.exists(result_file_dir):
        os.makedirs(result_file_dir)
    (r1, r5, r10, medr, meanr, mir, mAP) = result_tuple
    text = " * %s:\n" % name_str
    text += " * r_1_5_10: {}\n".format([round(r1, 3), round(r5, 3), round(r10, 3)])
    text += " * medr, meanr, mir: {}\n".format([round(medr, 3), round(meanr, 3), round(mir, 3)])
    text += " * mAP: {}\n".format(round(mAP, 3))
    text += " * " + "-" * 10
    print(text)
    opt = None if checkpoint is None else (checkpoint["opt"] if isinstance(checkpoint, Mapping) else checkpoint.opt)
    with open(predict_result_file, "a") as f:
        f.write(str(time.asctime(time.localtime(time.time()))) + "\t")
        for each in [model_path, round(r1, 3), round(r5, 3), round(r10, 3), round(medr, 3), round(meanr, 3), round(mir, 3),
                     round(mAP, 3)]:
            f.write(str(each))
            f.write("\t")
        f.write(opt.parm_adjust_config.replace("_", "\t") if opt is not None else "")
        f.write("\n")


# ----------------------------------------------------------------------------------------------------------------
# the two callers
# ----------------------------------------------------------------------------------------------------------------
def evaluate_and_write(t2i_matrix, txt_ids, vis_ids, output_dir, predict_result_file, model_path, checkpoint,
                       captions: Optional[Mapping[str, str]] = None, txt_loader=None, pred_result_file=None,
                       with_ground_truth: bool = True):
    """The tail of predictor.get_predict_file (predictor.py:232-284) for one (query set, collection) pair.

    with_ground_truth (collections with labelled captions): metrics in both directions appended to
    <dir>/TextToVideo/<file> and <dir>/VideoToText/<file>, plus <output_dir>/t2v.pkl (top 500).  Otherwise (ad-hoc
    queries): t2v.pkl and the id.sent.score.txt list (top 2000).  Returns a dict of what was computed."""
    from .retrieval import RankedScores
    if isinstance(t2i_matrix, RankedScores):     # large gallery (W2VVPP.predict_batch): no Q x V matrix anywhere
        return evaluate_and_write_ranked(t2i_matrix, txt_ids, vis_ids, output_dir, predict_result_file, model_path, checkpoint,
                                         captions=captions, txt_loader=txt_loader, pred_result_file=pred_result_file,
                                         with_ground_truth=with_ground_truth)
    s = _as_scores(t2i_matrix)
    os.makedirs(output_dir, exist_ok=True)
    out = {}
    if with_ground_truth:
        t2v, rank0 = evaluate_t2v(s, txt_ids, vis_ids)
        result_dir, result_name = os.path.dirname(predict_result_file), os.path.basename(predict_result_file)
        write_to_predict_result_file(os.path.join(result_dir, "TextToVideo", result_name), model_path, checkpoint, t2v)
        txt2video_write_to_file(None, None, vis_ids, txt_ids, s, pkl_saved_file=os.path.join(output_dir, "t2v.pkl"),
                                txt_loader=txt_loader, Threshold=500, captions=captions)
        v2t, _, _ = evaluate_v2t(s, txt_ids, vis_ids)
        write_to_predict_result_file(os.path.join(result_dir, "VideoToText", result_name), model_path, checkpoint, v2t,
                                     name_str="Video To Text")
        out.update(t2v=t2v, v2t=v2t, rank0=rank0)
        return out
    txt2video_write_to_file(None, None, vis_ids, txt_ids, s, pkl_saved_file=os.path.join(output_dir, "t2v.pkl"),
                            txt_loader=txt_loader, Threshold=500, captions=captions)
    if pred_result_file is None:
        pred_result_file = os.path.join(output_dir, "id.sent.score.txt")
    txt2video_write_to_file(pred_result_file, None, vis_ids, txt_ids, s)
    out.update(pred_result_file=pred_result_file)
    return out


def evaluate_and_write_ranked(pred, txt_ids, vis_ids, output_dir, predict_result_file, model_path, checkpoint,
                              captions: Optional[Mapping[str, str]] = None, txt_loader=None, pred_result_file=None,
                              with_ground_truth: bool = True):
    """evaluate_and_write for a retrieval.RankedScores (gallery too large for a dense score matrix): text->video ranks,
    top-10 and metrics from ONE fused similarity sweep (laff_sim_rank_topk), the written lists (t2v.pkl top 500,
    id.sent.score.txt top 2000) from the threshold sweep (laff_sim_collect) -- the files and numbers are those of the
    dense path (tests/test_gpu_collection.py, test_gpu_predictor.py).  The video->text direction needs every column of
    the matrix and is only evaluated on the dense path."""
    os.makedirs(output_dir, exist_ok=True)
    out = {}
    dev = pred.q16.device
    k500 = writer_topk(len(vis_ids), 500)
    if with_ground_truth:
        gt = torch.from_numpy(gt_index(txt_ids, vis_ids)).to(dev)
        # ONE sweep: the candidates of the top-500 lists and the exact rank of every ground truth (laff_sim_collect_rank)
        vals, idx, rank0 = pred.ranked_lists(max(k500, 1), gt_global=gt)
        out["t2v"] = _metrics_tuple(ops.rank_metrics(rank0))
        out["rank0"] = rank0
        names = ("r1", "r5", "r10", "medr", "meanr", "mir", "mAP")
        m = dict(zip(names, out["t2v"]))
        print(" * Text to video:")
        print(" * r_1_5_10: {}".format([round(m["r1"], 3), round(m["r5"], 3), round(m["r10"], 3)]))
        print(" * medr, meanr, mir: {}".format([round(m["medr"], 3), round(m["meanr"], 3), round(m["mir"], 3)]))
        write_to_predict_result_file(os.path.join(os.path.dirname(predict_result_file), "TextToVideo", os.path.basename(predict_result_file)),
                                     model_path, checkpoint, out["t2v"])
    else:
        vals, idx = pred.ranked_lists(max(k500, writer_topk(len(vis_ids), 2000)))
    vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
    txt2video_write_to_file(None, (vals, idx), vis_ids, txt_ids, None, pkl_saved_file=os.path.join(output_dir, "t2v.pkl"),
                            txt_loader=txt_loader, Threshold=500, captions=captions)
    if not with_ground_truth:
        f = pred_result_file or os.path.join(output_dir, "id.sent.score.txt")
        txt2video_write_to_file(f, (vals, idx), vis_ids, txt_ids, None)
        out["pred_result_file"] = f
    return out


def validate(model, txt_loader, vis_loader, epoch=None, measure="cosine", metric="mir", negative_val=False, config=None,
             negation_set: Optional[Sequence[str]] = None):
    """trainer.validate (trainer.py:579-607): text->video metrics of the current model on the validation loaders.
    Returns (the metric named by `metric` (default mir), mir over the negation subset or None)."""
    measure = getattr(config, "measure", measure) if config is not None else measure
    s, txt_ids, vis_ids = model.predict_device(txt_loader, vis_loader, measure)
    t2v, rank0 = evaluate_t2v(s, txt_ids, vis_ids)
    names = ("r1", "r5", "r10", "medr", "meanr", "mir", "mAP")
    vals = dict(zip(names, t2v))
    print(" * Text to video:")
    print(" * r_1_5_10: {}".format([round(vals["r1"], 3), round(vals["r5"], 3), round(vals["r10"], 3)]))
    print(" * medr, meanr, mir: {}".format([round(vals["medr"], 3), round(vals["meanr"], 3), round(vals["mir"], 3)]))
    print(" * mAP: {}".format(round(vals["mAP"], 3)))
    mir2 = None
    if negative_val:
        neg = set(negation_set or ())
        sel = [i for i, t in enumerate(txt_ids) if t in neg]
        if sel:
            sub = rank0[torch.as_tensor(sel, device=rank0.device)]
            mir2 = _metrics_tuple(ops.rank_metrics(sub))[5]
    return vals.get(metric, vals["mir"]), mir2
