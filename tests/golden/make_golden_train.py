"""Golden vectors for the training step from the UNMODIFIED reference: the 'LAFF' model
(model.model.W2VVPP_MultiHeadAttention, configs.laff) in train mode, three calls of `model(train_data, epoch)` —
forward, summed per-head MarginRankingLoss, backward, clip_grad_norm_(params, 2), RMSprop / Adam step.

    python tests/golden/make_golden_train.py     # writes tests/golden/train_<case>.npz

Dropout is set to 0 (torch's dropout mask cannot be reproduced outside torch); everything else is the shipped setting.
Text encoders are the pass-through stand-ins of make_golden.py (their features are inputs at this tier).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from laff_b200 import synth  # noqa: E402

SMALL = mg.SMALL


def run_case(mm, tag, optimizer, lr, batch_norm, steps=3, B=16, D=256, H=8, seed=81, with_ave=False, mul=False, loss="mrl"):
    import torch
    dims = SMALL
    vis_dims = {synth.VIS_CLIP_FT: dims["clip"], synth.VIS_TF: dims["tf"], synth.VIS_X3D: dims["x3d"], synth.VIS_IRCSN: dims["ircsn"]}
    cfg = mg.make_config("laff", D, H, vis_dims, dims, with_ave, mul)
    cfg.dropout = 0.0
    cfg.batch_norm = batch_norm
    cfg.optimizer, cfg.lr = optimizer, lr
    cfg.loss = loss
    torch.manual_seed(0)
    model = mm.W2VVPP_MultiHeadAttention(cfg)
    sd0 = mg.load_synth_state(model, seed, omega=0.6 if with_ave else 1.0)
    model.train()
    out = {"meta": np.array([B, D, H, steps, seed, int(batch_norm), 0, int(with_ave), int(mul)]), "loss_kind": np.array(loss),
           "optimizer": np.array(optimizer), "lr": np.float64(lr),
           "grad_clip": np.float64(cfg.grad_clip), "vis_names": np.array(list(vis_dims.keys()))}
    for k, v in sd0.items():
        out["sd0/" + k] = v
    losses = []
    for s in range(steps):
        vis_in = {}
        for name, d in vis_dims.items():
            vis_in[name] = synth.feature(seed + s, "vis/" + name, B, d, "dense" if name == synth.VIS_CLIP_FT else "relu")
        txt_in = {"gru": synth.feature(seed + s, "txt/gru", B, dims["gru"]), "bow": synth.feature(seed + s, "txt/bow", B, dims["bow"], "bow"),
                  "w2v": synth.feature(seed + s, "txt/w2v", B, dims["w2v"]), "clip": synth.feature(seed + s, "txt/clip", B, dims["clip"])}
        for k, v in vis_in.items():
            out["step%d/vin/%s" % (s, k)] = v
        for k, v in txt_in.items():
            out["step%d/tin/%s" % (s, k)] = v
        train_data = {"vis_feats": {k: torch.from_numpy(v) for k, v in vis_in.items()},
                      "captions": {k: torch.from_numpy(v) for k, v in txt_in.items()},
                      "captions_task2": None, "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}
        items = model(train_data, epoch=0)
        losses.append(float(items["triplet_loss"]))
        if s == 0:  # gradients of the first step (after clipping, as they sit in .grad)
            for k, p in model.named_parameters():
                if p.grad is not None:
                    out["grad0/" + k] = p.grad.detach().numpy().copy()
        for k, v in model.state_dict().items():
            out["sd%d/%s" % (s + 1, k)] = v.detach().numpy().copy()
    out["losses"] = np.array(losses)
    print(tag, "losses", losses)
    return out


def run_gru_case(mm, tag, optimizer, lr, steps=3, B=12, D=256, H=8, seed=95):
    """'LAFF' with the reference's real GruTxtEncoder (trainable embedding + GRU, backward through time by autograd) on
    the synthetic vocabulary of tests/golden/text; BoW / word2vec / CLIP features stay pass-through inputs."""
    import json
    import torch
    import txt2vec
    dims = dict(SMALL, gru=24)
    vis_dims = {synth.VIS_CLIP_FT: dims["clip"], synth.VIS_TF: dims["tf"], synth.VIS_X3D: dims["x3d"], synth.VIS_IRCSN: dims["ircsn"]}
    cfg = mg.make_config("laff", D, H, vis_dims, dims)
    cfg.dropout = 0.0
    cfg.optimizer, cfg.lr = optimizer, lr
    cfg.t2v_idx = txt2vec.IndexVec(os.path.join(HERE, "text", "vocab_gru.pkl"))
    cfg.we_dim, cfg.rnn_layer = 12, 1
    cfg.we = torch.zeros(len(cfg.t2v_idx.vocab), cfg.we_dim)
    real_gru = mg.REAL_GRU_ENCODER
    saved = mm.GruTxtEncoder
    mm.GruTxtEncoder = real_gru
    try:
        torch.manual_seed(0)
        model = mm.W2VVPP_MultiHeadAttention(cfg)
    finally:
        mm.GruTxtEncoder = saved
    sd0 = mg.load_synth_state(model, seed)
    model.train()
    meta = json.load(open(os.path.join(HERE, "text", "meta.json")))
    words = meta["bow_words"]
    out = {"meta": np.array([B, D, H, steps, seed, 0, 0, 0, 0]), "loss_kind": np.array("mrl"), "optimizer": np.array(optimizer),
           "lr": np.float64(lr), "grad_clip": np.float64(cfg.grad_clip), "vis_names": np.array(list(vis_dims.keys())),
           "gru_dim": np.int64(dims["gru"]), "we_dim": np.int64(cfg.we_dim)}
    for k, v in sd0.items():
        out["sd0/" + k] = v
    losses = []
    for s in range(steps):
        r = synth.rng_for(seed + s, "captions")
        caps = [" ".join(r.choice(words + ["zebra", "the"], size=r.randint(1, 9))) for _ in range(B)]
        vis_in = {n: synth.feature(seed + s, "vis/" + n, B, d, "dense" if n == synth.VIS_CLIP_FT else "relu") for n, d in vis_dims.items()}
        txt_in = {"bow": synth.feature(seed + s, "txt/bow", B, dims["bow"], "bow"), "w2v": synth.feature(seed + s, "txt/w2v", B, dims["w2v"]),
                  "clip": synth.feature(seed + s, "txt/clip", B, dims["clip"])}
        for k, v in vis_in.items():
            out["step%d/vin/%s" % (s, k)] = v
        for k, v in txt_in.items():
            out["step%d/tin/%s" % (s, k)] = v
        out["step%d/captions" % s] = np.array(caps)
        captions = {k: torch.from_numpy(v) for k, v in txt_in.items()}
        captions["caption"] = caps
        train_data = {"vis_feats": {k: torch.from_numpy(v) for k, v in vis_in.items()}, "captions": captions, "captions_task2": None,
                      "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}
        items = model(train_data, epoch=0)
        losses.append(float(items["triplet_loss"]))
        if s == 0:
            for k, p in model.named_parameters():
                if p.grad is not None:
                    out["grad0/" + k] = p.grad.detach().numpy().copy()
        for k, v in model.state_dict().items():
            out["sd%d/%s" % (s + 1, k)] = v.detach().numpy().copy()
    out["losses"] = np.array(losses)
    print(tag, "losses", losses)
    return out


def run_frame_case(mm, tag, optimizer, lr, steps=3, B=16, D=256, H=8, F=6, seed=91, amp=False):
    """'FrameLAFF' (LAFF-ml, W2VVPP_MutiVisFrameFeat): frame-level attention + BatchNorm on every projected feature.

    amp=True runs the reference's float16 branch (model/model.py:970-989: autocast, scaler.scale(loss).backward(),
    clip_grad_norm_ on the scaled gradients, scaler.step / update) -- the branch configs/FrameLaff_...:33 selects.  The
    reference imports torch.cuda.amp's autocast / GradScaler, which switch themselves off without a CUDA device; here
    the two names are bound to torch's CPU flavours of the same classes (fp16 autocast, GradScaler('cpu')), so the
    reference's own step code runs with a live loss scale.  Per step the loss, the scale after scaler.update() and
    whether the step was skipped are recorded; state dicts only after the first executed step and at the end."""
    import torch
    if amp:
        mm.autocast = lambda: torch.autocast("cpu", dtype=torch.float16)
        mm.GradScaler = lambda: torch.amp.GradScaler("cpu")
        mm.float16 = True
    dims = SMALL
    vis_dims = {synth.VIS_C3D: dims["c3d"], synth.VIS_TF: dims["tf"], synth.VIS_X3D: dims["x3d"], synth.VIS_IRCSN: dims["ircsn"],
                synth.VIS_FRAME: dims["clip"]}
    cfg = mg.make_config("frame", D, H, vis_dims, dims)
    cfg.dropout = 0.0
    cfg.float16 = bool(amp)        # False: the fp32 path is the numerical reference
    cfg.optimizer, cfg.lr = optimizer, lr
    torch.manual_seed(0)
    model = mm.W2VVPP_MutiVisFrameFeat(cfg)
    sd0 = mg.load_synth_state(model, seed)
    model.train()
    out = {"meta": np.array([B, D, H, steps, seed, 1, F]), "optimizer": np.array(optimizer), "lr": np.float64(lr),
           "grad_clip": np.float64(cfg.grad_clip), "vis_names": np.array([n for n in vis_dims if n != synth.VIS_FRAME]),
           "frame_feat": np.array(synth.VIS_FRAME)}
    for k, v in sd0.items():
        out["sd0/" + k] = v
    losses, scales, skipped, first_done = [], [], [], None
    for s in range(steps):
        vis_in = {n: synth.feature(seed + s, "vis/" + n, B, d, "relu") for n, d in vis_dims.items() if n != synth.VIS_FRAME}
        fr = synth.feature(seed + s, "frames", B * F, dims["clip"]).reshape(B, F, dims["clip"])
        lens = synth.rng_for(seed + s, "lens").randint(1, F + 1, size=B)
        lens[0] = F
        mask = np.ones((B, F), dtype=np.float32)
        for i, n in enumerate(lens):  # zero-padded tails as collate_pair produces them (data_provider.py:108-121)
            fr[i, n:] = 0
            mask[i, n:] = 0
        txt_in = {"gru": synth.feature(seed + s, "txt/gru", B, dims["gru"]), "bow": synth.feature(seed + s, "txt/bow", B, dims["bow"], "bow"),
                  "w2v": synth.feature(seed + s, "txt/w2v", B, dims["w2v"]), "clip": synth.feature(seed + s, "txt/clip", B, dims["clip"])}
        for k, v in vis_in.items():
            out["step%d/vin/%s" % (s, k)] = v
        for k, v in txt_in.items():
            out["step%d/tin/%s" % (s, k)] = v
        out["step%d/frames" % s], out["step%d/mask" % s] = fr.copy(), mask
        train_data = {"vis_feats": {k: torch.from_numpy(v) for k, v in vis_in.items()},
                      "captions": {k: torch.from_numpy(v) for k, v in txt_in.items()}, "captions_task2": None,
                      "vis_frame_feat_dict": {"mask_tensor": torch.from_numpy(mask), synth.VIS_FRAME: torch.from_numpy(fr.copy())},
                      "vis_origin_frame_tuple": None}
        if amp:
            before = model.scaler.get_scale()
        items = model(train_data, epoch=0)
        losses.append(float(items["triplet_loss"].detach()))
        if amp:
            scales.append(model.scaler.get_scale())
            skipped.append(int(scales[-1] < before))
            if not skipped[-1] and first_done is None:
                first_done = s
                out["first_executed_step"] = np.int64(s)
                for k, p in model.named_parameters():   # unscaled + clipped, as scaler.step() leaves them in .grad
                    if p.grad is not None:
                        out["grad_first/" + k] = p.grad.detach().float().numpy().copy()
            if s == first_done or s == steps - 1:
                for k, v in model.state_dict().items():
                    out["sd%d/%s" % (s + 1, k)] = v.detach().float().numpy().copy()
            continue
        if s == 0:
            for k, p in model.named_parameters():
                if p.grad is not None:
                    out["grad0/" + k] = p.grad.detach().numpy().copy()
        for k, v in model.state_dict().items():
            out["sd%d/%s" % (s + 1, k)] = v.detach().numpy().copy()
    out["losses"] = np.array(losses)
    if amp:
        out["scales"], out["skipped"] = np.array(scales), np.array(skipped)
        mm.float16 = False
        print(tag, "scales", scales, "skipped", skipped)
    print(tag, "losses", losses)
    return out


def main():
    import torch
    torch.set_num_threads(8)
    mm, rloss, reval, ratt = mg.import_reference()
    only = sys.argv[1:]
    for tag, optimizer, lr, bn in (("rmsprop", "rmsprop", 1e-3, False), ("adam", "adam", 1e-3, False), ("rmsprop_bn", "rmsprop", 1e-3, True)):
        if not only or tag in only:
            np.savez_compressed(os.path.join(HERE, "train_%s.npz" % tag), **run_case(mm, tag, optimizer, lr, bn))
    if not only or "rmsprop_ave_mul" in only:
        np.savez_compressed(os.path.join(HERE, "train_rmsprop_ave_mul.npz"),
                            **run_case(mm, "rmsprop_ave_mul", "rmsprop", 1e-3, False, with_ave=True, mul=True))
    if not only or "adam_dsl" in only:
        np.savez_compressed(os.path.join(HERE, "train_adam_dsl.npz"), **run_case(mm, "adam_dsl", "adam", 1e-3, False, loss="dsl"))
    if not only or "gru_rmsprop" in only:
        np.savez_compressed(os.path.join(HERE, "train_gru_rmsprop.npz"), **run_gru_case(mm, "gru_rmsprop", "rmsprop", 1e-3))
    if not only or "frame_rmsprop" in only:
        np.savez_compressed(os.path.join(HERE, "train_frame_rmsprop.npz"), **run_frame_case(mm, "frame_rmsprop", "rmsprop", 1e-3))
    for tag, optimizer in (("frame_amp_rmsprop", "rmsprop"), ("frame_amp_adam", "adam")):
        if not only or tag in only:
            np.savez_compressed(os.path.join(HERE, "train_%s.npz" % tag),
                                **run_frame_case(mm, tag, optimizer, 1e-3, steps=12, amp=True))


if __name__ == "__main__":
    main()
