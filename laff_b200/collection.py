"""Text-to-video prediction over an on-disk collection in the reference's layout — the body of
`predictor.get_predict_file` (predictor.py:129-284) after the checkpoint is loaded, end to end on the device:

    <rootpath>/<collection>/FeatureData/<feature>/{feature.bin,id.txt,shape.txt}     video-level features (BigFile)
    <rootpath>/<collection>/VideoSets/<collection>.txt                                gallery video ids, one per line
    <rootpath>/<collection>/TextData/<query_set>                                      "<cap_id> <caption>" per line
    <rootpath>/<collection>/TextData/<dir_name>/...                                   precomputed text features (BigFile
                                                                                     keyed by cap_id), for encodings whose
                                                                                     config entry has a 'dir_name'
    -> <rootpath>/<collection>/SimilarityIndex/<query_set>/<sim_name>/{t2v.pkl, id.sent.score.txt}
       <predict_result_file dir>/{TextToVideo,VideoToText}/<file>                      appended metric lines

The reference walks the gallery and the queries through DataLoaders item by item (one file seek per video per
feature), keeps the gallery embeddings on the host, re-uploads them per text batch, argsorts on the host.  Here the
feature files are streamed to the GPU shard by shard (`bigfile.load_features`), the gallery is fused once into a
resident 16-bit index, queries are encoded from strings + precomputed features on the device, and ranking / lists /
metrics come from `laff_b200.predictor`.
"""
from __future__ import annotations

import os
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch

from . import predictor as _pred
from .bigfile import BigFile, load_features
from .retrieval import GalleryIndex, Retriever

NO_GROUND_TRUTH_COLLECTIONS = ("iacc.3", "v3c1")  # predictor.py:234


def read_captions(capfile: str) -> Tuple[List[str], Dict[str, str]]:
    """TextDataset's caption file parsing (data_provider.py:551-561): '<cap_id> <caption>'; a line with only an id has
    an empty caption; blank lines are skipped; a repeated id keeps its last caption but is listed every time."""
    cap_ids, captions = [], {}
    with open(capfile, "r") as reader:
        for line in reader.readlines():
            if line.strip() == "":
                continue
            parts = line.strip().split(None, 1)
            cap_id, caption = (parts[0], "") if len(parts) < 2 else parts
            captions[cap_id] = caption
            cap_ids.append(cap_id)
    return cap_ids, captions


def precalculated_text_features(config, text_dir: str) -> Dict[str, BigFile]:
    """TextDataset.get_precalculate_file (data_provider.py:564-573)."""
    out = {}
    for name, enc in config.text_encoding.items():
        if "no" in enc["name"]:
            continue
        if "dir_name" in enc and enc["dir_name"]:
            out[name] = BigFile(os.path.join(text_dir, enc["dir_name"]))
    return out


def caption_features(config, text_dir: str, cap_ids: Sequence[str], captions: Mapping[str, str], device) -> dict:
    """The caption_feat_dict of a whole query set: {'caption': [str], '<encoding>': device tensor [Q, d], ...}."""
    feats: dict = {"caption": [captions[c] for c in cap_ids]}
    for name, bf in precalculated_text_features(config, text_dir).items():
        feats[name] = load_features({name: bf}, list(cap_ids), device)[name]
    return feats


def build_gallery(model, config, rootpath: str, collection: str, device, rank: int = 0, world_size: int = 1):
    """Gallery ids + the resident index of this rank's shard (predictor.py:190-214 + the vis loop of model.predict)."""
    # base_config.py:171-173 ships a non-empty vid_frame_feats together with frame_feat_input = False: only the flag
    # decides whether frame files are read (predictor.py:191)
    if getattr(config, "frame_feat_input", False):
        raise NotImplementedError("frame-level feature files (FeatureData/frame) are not read yet: LAFF collections only")
    with open(os.path.join(rootpath, collection, "VideoSets", collection + ".txt")) as f:
        vis_ids = list(map(str.strip, f))
    from .retrieval import shard_bounds
    lo, hi = shard_bounds(len(vis_ids), world_size, rank)
    files = {y: BigFile(os.path.join(rootpath, collection, "FeatureData", y)) for y in config.vid_feats}
    feats = load_features(files, vis_ids[lo:hi], device)
    index = GalleryIndex.from_features(model.vis_net, feats, len(vis_ids), rank, world_size)
    return vis_ids, index


def predict_collection(model, config, rootpath: str, collection: str, query_sets: Sequence[str], sim_name: str,
                       predict_result_file: Optional[str] = None, model_path: str = "", checkpoint=None, device=None,
                       dense_limit: int = 1 << 28):
    """Runs every query set of `collection` (see the module docstring).  Returns {query_set: result dict}.
    Query sets with ground truth ('<vid>#...' caption ids; not iacc.3 / v3c1 / simple_query.txt, predictor.py:234) get
    metrics in both directions + t2v.pkl; the others t2v.pkl + id.sent.score.txt.  Galleries whose Q x V score matrix
    would exceed `dense_limit` entries use the fused sweep for ranks and chunked dense lists for the files."""
    device = torch.device(device) if device is not None else next(model.parameters()).device
    model.eval()
    vis_ids, index = build_gallery(model, config, rootpath, collection, device)
    retr = Retriever(model.txt_net, index)
    results = {}
    for query_set in query_sets:
        output_dir = os.path.join(rootpath, collection, "SimilarityIndex", query_set, sim_name)
        os.makedirs(output_dir, exist_ok=True)
        text_dir = os.path.join(rootpath, collection, "TextData")
        cap_ids, captions = read_captions(os.path.join(text_dir, query_set))
        feats = caption_features(config, text_dir, cap_ids, captions, device)
        with_gt = collection not in NO_GROUND_TRUTH_COLLECTIONS and query_set != "simple_query.txt"
        q16 = retr.encode_queries(feats)
        if len(cap_ids) * len(vis_ids) <= dense_limit:
            from . import ops
            scores = ops.sim_dense(q16, index.g16, 1.0 / index.heads)
            results[query_set] = _pred.evaluate_and_write(
                scores, cap_ids, vis_ids, output_dir, predict_result_file or os.path.join(output_dir, "predict_result.txt"),
                model_path + "\t" + collection, checkpoint, captions=captions, with_ground_truth=with_gt)
            continue
        # large gallery: no Q x V matrix
        out = {}
        if with_gt:
            gt = torch.from_numpy(_pred.gt_index(cap_ids, vis_ids)).to(device)
            res = index.search(q16, gt, 10)
            out["t2v"] = _pred._metrics_tuple(res.metrics)
            out["rank0"] = res.rank0
            prf = predict_result_file or os.path.join(output_dir, "predict_result.txt")
            _pred.write_to_predict_result_file(os.path.join(os.path.dirname(prf), "TextToVideo", os.path.basename(prf)),
                                               model_path + "\t" + collection, checkpoint, out["t2v"])
        k500 = _pred.writer_topk(len(vis_ids), 500)
        vals, idx = index.ranked_lists(q16, max(k500, 1 if with_gt else _pred.writer_topk(len(vis_ids), 2000)))
        vals, idx = vals.cpu().numpy(), idx.cpu().numpy()
        _pred.txt2video_write_to_file(None, (vals, idx), vis_ids, cap_ids, None, pkl_saved_file=os.path.join(output_dir, "t2v.pkl"),
                                      Threshold=500, captions=captions)
        if not with_gt:
            f = os.path.join(output_dir, "id.sent.score.txt")
            _pred.txt2video_write_to_file(f, (vals, idx), vis_ids, cap_ids, None)
            out["pred_result_file"] = f
        results[query_set] = out
    return results
