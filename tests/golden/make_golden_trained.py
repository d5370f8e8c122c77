"""The trained-checkpoint fixture (SURVEY §8c/§8d, T2 end-to-end parity): the UNMODIFIED reference 'LAFF' model
(model.model.W2VVPP_MultiHeadAttention, configs.laff) trained on CPU by its own step code -- `model(train_data, epoch)`
(model/model.py:964-1001), summed per-head MarginRankingLoss (loss.py:95-135), clip_grad_norm_, RMSprop -- on a
synthetic latent-factor collection (laff_b200.synth.latent_collection), then evaluated by the reference's own
evaluation path on held-out collections of the C1 (1000 x 1000) and C2 (2990 x 2990) sizes:
vis_net / txt_net (eval) -> get_txt2vis_matrix -> np.argsort (predictor.py:232) -> evaluation.eval_qry2retro.

    python tests/golden/make_golden_trained.py        # writes tests/golden/trained_laff.npz (~4 MB)

Stored: the trained state dict, the training-loss curve, and per evaluation size the reference's rank of the ground
truth of every query (read off its argsort exactly as predictor.py:239-244 does), its top-10 lists, R@1/5/10, MedR,
the score of the ground truth and of the 11 best videos, the smallest score gap at any rank <= 11 and the distance of
the ground truth's score to its nearest competitor (the margins a pipeline has to resolve to return the same lists /
ranks: queries whose margin lies below the fp32 accumulation noise are enumerated by the test, not hidden).  The evaluation features are not
stored: synth.latent_collection regenerates them from the seeds recorded here.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from laff_b200 import synth  # noqa: E402

TRAIN_SEED, TRAIN_PAIRS, BATCH, STEPS = 300, 6000, 128, 300
NOISE = dict(vis_noise=float(os.environ.get("LAFF_TRAINED_VN", 1.0)), cap_noise=float(os.environ.get("LAFF_TRAINED_CN", 1.2)),
             txt_noise=float(os.environ.get("LAFF_TRAINED_TN", 1.0)))
EVAL = {"c1": (311, 1000), "c2": (312, 2990)}


def main():
    import torch
    torch.set_num_threads(max(1, (os.cpu_count() or 2)))
    mm, rloss, reval, ratt = mg.import_reference()
    dims, D, H = synth.TRAINED_DIMS, synth.TRAINED_D, synth.TRAINED_HEADS
    vis_dims = {synth.VIS_CLIP_FT: dims["clip"], synth.VIS_TF: dims["tf"], synth.VIS_X3D: dims["x3d"], synth.VIS_IRCSN: dims["ircsn"]}
    cfg = mg.make_config("laff", D, H, vis_dims, dims)
    lr = float(os.environ.get("LAFF_TRAINED_LR", cfg.lr))
    cfg.lr = lr
    torch.manual_seed(0)
    model = mm.W2VVPP_MultiHeadAttention(cfg)      # shipped settings: dropout 0.2, tanh, rmsprop, grad_clip 2, margin 0.2
    model.train()
    vis_tr, txt_tr = synth.latent_collection(TRAIN_SEED, TRAIN_PAIRS, **NOISE)
    pick = np.random.RandomState(7)
    losses = []
    for step in range(STEPS):
        b = pick.choice(TRAIN_PAIRS, BATCH, replace=False)
        train_data = {"vis_feats": {k: torch.from_numpy(v[b]) for k, v in vis_tr.items()},
                      "captions": {k: torch.from_numpy(v[b]) for k, v in txt_tr.items()},
                      "captions_task2": None, "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}
        losses.append(float(model(train_data, epoch=0)["triplet_loss"].detach()))
        if step % 50 == 0 or step == STEPS - 1:
            print("step %d loss %.3f" % (step, losses[-1]), flush=True)
    model.eval()
    out = {"meta": np.array([D, H, STEPS, BATCH, TRAIN_SEED, TRAIN_PAIRS]),
           "noise": np.array([NOISE["vis_noise"], NOISE["cap_noise"], NOISE["txt_noise"]]), "lr": np.float64(lr), "losses": np.array(losses, dtype=np.float32),
           "vis_names": np.array(list(vis_dims.keys()))}
    for k, v in model.state_dict().items():
        out["sd/" + k] = v.detach().numpy().copy()
    for tag, (seed, n) in EVAL.items():
        vis, txt = synth.latent_collection(seed, n, **NOISE)
        with torch.no_grad():
            v_emb = model.vis_net({k: torch.from_numpy(x) for k, x in vis.items()})
            t_emb = model.txt_net({k: torch.from_numpy(x) for k, x in txt.items()})
            scores = model.get_txt2vis_matrix(t_emb, v_emb).numpy()
        inds = np.argsort(scores, axis=1)                       # predictor.py:232
        rank0 = np.empty(n, dtype=np.int64)
        for i in range(n):                                      # predictor.py:239-244 with integer ids
            rank0[i] = np.where(inds[i][::-1] == i)[0][0]
        r1, r5, r10, medr, meanr, mir = reval.eval_qry2retro(scores, n_qry=1)
        top = inds[:, ::-1][:, :11]
        tops = np.take_along_axis(scores, top, 1)
        srt = -np.sort(-scores, axis=1)[:, :12]
        gaps = srt[:, :-1] - srt[:, 1:]                         # gaps between consecutive ranks 1..12
        sg = scores[np.arange(n), np.arange(n)]
        diff = np.abs(scores - sg[:, None])
        diff[np.arange(n), np.arange(n)] = np.inf
        out[tag + "/gt_gap"] = diff.min(1).astype(np.float32)       # distance of the nearest competitor to the ground truth's score
        out.update({tag + "/seed": np.int64(seed), tag + "/n": np.int64(n), tag + "/rank0": rank0, tag + "/top10": top[:, :10].astype(np.int32),
                    tag + "/top_scores": tops.astype(np.float32), tag + "/s_gt": scores[np.arange(n), np.arange(n)].astype(np.float32),
                    tag + "/metrics": np.array([r1, r5, r10, medr, meanr, mir], dtype=np.float64),
                    tag + "/min_gap_top11": gaps[:, :11].min(1).astype(np.float32),
                    tag + "/emb_txt_sample": t_emb.numpy()[:8].copy(), tag + "/emb_vis_sample": v_emb.numpy()[:8].copy()})
        print(tag, "R@1 %.2f R@5 %.2f R@10 %.2f MedR %.0f  | min top-11 gap: median %.2e min %.2e" %
              (r1, r5, r10, medr, float(np.median(gaps[:, :11].min(1))), float(gaps[:, :11].min())))
    np.savez_compressed(os.path.join(HERE, "trained_laff.npz"), **out)


if __name__ == "__main__":
    main()
