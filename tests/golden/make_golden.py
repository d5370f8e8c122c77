"""Generate golden vectors by running the UNMODIFIED reference (ruc-aimc-lab/LAFF at /root/reference) on CPU.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference ships no golden vectors or known-answer tests, so parity is pinned by outputs of the
reference's own classes: configs.laff / configs.FrameLaff... config objects, VisMutiTransformNetAddAttnetion,
MultiScaleTxtEncoderAttention, VisMutiTransformNetPlusFrameFeat, Multi_head_MyApply_Attention, Attention_1,
W2VVPP.get_txt2vis_matrix, loss.MarginRankingLoss(+WithScore), evaluation.eval_qry2retro / eval / cosine_sim.
Five leaf imports that are absent here are shimmed (ftfy, nltk, prefetch_generator, torchvision Kinetics400) and the
four text encoders are replaced by pass-through modules (the text features are inputs at this tier); everything
downstream is the reference's code.  This script only runs where /root/reference exists; the .npz files it writes
are committed and are what the tests read.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("LAFF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from laff_b200 import synth  # noqa: E402


REAL_GRU_ENCODER = None


def install_shims():
    os.environ.setdefault("HOME", "/tmp")
    for name in ("ftfy", "prefetch_generator"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.fix_text = lambda s: s
            m.BackgroundGenerator = object
            sys.modules[name] = m
    if "nltk" not in sys.modules:
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
    import torchvision.datasets as tvd
    if not hasattr(tvd, "Kinetics400"):
        tvd.Kinetics400 = getattr(tvd, "Kinetics", object)
    if REF not in sys.path:
        sys.path.insert(0, REF)


def import_reference():
    install_shims()
    import torch
    import model.model as mm  # noqa  (reference)
    import loss as rloss
    import evaluation as reval
    import model.Attention as ratt
    mm.device = torch.device("cpu")
    mm.float16 = False

    class PassThrough(torch.nn.Module):
        """Stands in for GruTxtEncoder / BoWTxtEncoder / W2VTxtEncoder / CLIPEncoder: returns the given feature."""
        key = None

        def __init__(self, opt=None):
            super().__init__()

        def forward(self, caption_feat_dict, task3=False):
            return {"text_features": caption_feat_dict[self.key]}

    def pt(key):
        return type("PassThrough_" + key, (PassThrough,), {"key": key})

    global REAL_GRU_ENCODER
    REAL_GRU_ENCODER = getattr(mm.GruTxtEncoder, "_laff_real", mm.GruTxtEncoder)   # the reference's own class (make_golden_train)
    mm.GruTxtEncoder = pt("gru")
    mm.GruTxtEncoder._laff_real = REAL_GRU_ENCODER
    mm.BoWTxtEncoder = pt("bow")
    mm.W2VTxtEncoder = pt("w2v")
    mm.CLIPEncoder = pt("clip")
    return mm, rloss, reval, ratt


def make_config(kind: str, D: int, H: int, vis_dims: dict, txt_dims: dict, with_ave=False, mul=False):
    """The reference's own config object, adjusted the way trainer.prepare_config does (trainer.py:130-135,156-157,212)."""
    import importlib
    if kind == "laff":
        cfg = importlib.import_module("configs.laff").config()
        cfg.adjust_parm("0_12_0_12_%d_%d_1" % (int(with_ave), int(mul)))
    else:
        cfg = importlib.import_module("configs.FrameLaff_NoFrameFc_StrongCLIP_adjust").config()
        cfg.adjust_parm("0_7_1_12_0_12_0")
        cfg.attention_param_each_head = {"with_ave": with_ave, "mul": mul, "split_head": True}
    cfg.vis_fc_layers = [dict(vis_dims), D]
    cfg.txt_fc_layers = [0, D]
    cfg.multi_head_attention = {"dropout": 0.0, "heads": H, "embed_dim_qkv": D // H}
    cfg.clip_opt = dict(cfg.clip_opt, size=txt_dims["clip"])
    cfg.rnn_size = txt_dims["gru"]
    cfg.t2v_bow = types.SimpleNamespace(ndims=txt_dims["bow"])
    cfg.t2v_w2v = types.SimpleNamespace(ndims=txt_dims["w2v"])
    return cfg


def load_synth_state(module, seed, omega=1.0):
    import torch
    sd = module.state_dict()
    new = {k: torch.from_numpy(np.asarray(synth.param(seed, k, tuple(v.shape), omega))).to(v.dtype).reshape(v.shape)
           for k, v in sd.items()}
    module.load_state_dict(new, strict=True)
    module.eval()
    return {k: v.numpy() for k, v in new.items()}


def run_fusion_case(mm, tag, D, H, dims, rows, seed, with_ave=False, mul=False, omega=1.0, bf16_inputs=False,
                    store_params=True):
    """LAFF config: vis_net + txt_net on synthetic features. Returns dict of arrays for the npz."""
    import torch
    vis_dims = {synth.VIS_CLIP_FT: dims["clip"], synth.VIS_TF: dims["tf"], synth.VIS_X3D: dims["x3d"],
                synth.VIS_IRCSN: dims["ircsn"]}
    cfg = make_config("laff", D, H, vis_dims, dims, with_ave, mul)
    vis_net = mm.VisMutiTransformNetAddAttnetion(cfg, cfg.vis_fc_layers[0])
    txt_net = mm.MultiScaleTxtEncoderAttention(cfg)
    out = {"meta": np.array([D, H, rows, seed, int(with_ave), int(mul), int(bf16_inputs)]), "omega": np.float32(omega)}
    vsd = load_synth_state(vis_net, seed, omega)
    tsd = load_synth_state(txt_net, seed + 1, omega)
    rnd = synth.bf16_round if bf16_inputs else (lambda a: a)
    if bf16_inputs:  # T1 tier: the tensor-core path sees bf16 operands; give the reference the same rounded values
        for sd, net in ((vsd, vis_net), (tsd, txt_net)):
            for k in sd:
                if k.endswith("fc1.weight"):
                    sd[k] = synth.bf16_round(sd[k])
            net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    vis_in = {}
    for name, d in vis_dims.items():
        kind = "dense" if name == synth.VIS_CLIP_FT else "relu"
        x = synth.feature(seed, "vis/" + name, rows, d, kind)
        vis_in[name] = rnd(x) if name != synth.VIS_CLIP_FT else x
    txt_in = {"gru": rnd(synth.feature(seed, "txt/gru", rows, dims["gru"])),
              "bow": synth.feature(seed, "txt/bow", rows, dims["bow"], "bow"),
              "w2v": rnd(synth.feature(seed, "txt/w2v", rows, dims["w2v"])),
              "clip": synth.feature(seed, "txt/clip", rows, dims["clip"])}
    with torch.no_grad():
        v_emb = vis_net({k: torch.from_numpy(v) for k, v in vis_in.items()})
        v_att = torch.stack([vis_net.attention_layer.attention_layer[h].weights for h in range(H)], 1)
        t_emb = txt_net({k: torch.from_numpy(v) for k, v in txt_in.items()})
        t_att = torch.stack([txt_net.attention_layer.attention_layer[h].weights for h in range(H)], 1)
    out["vis_emb"], out["txt_emb"] = v_emb.numpy(), t_emb.numpy()
    out["vis_att"], out["txt_att"] = v_att.numpy(), t_att.numpy()  # [rows, H, L]
    out["vis_names"] = np.array(list(vis_dims.keys()))
    out["vis_dims"] = np.array(list(vis_dims.values()))
    out["txt_dims"] = np.array([dims["gru"], dims["bow"], dims["w2v"], dims["clip"]])
    if store_params:
        for k, v in vis_in.items():
            out["vin/" + k] = v
        for k, v in txt_in.items():
            out["tin/" + k] = v
        for k, v in vsd.items():
            out["vsd/" + k] = v
        for k, v in tsd.items():
            out["tsd/" + k] = v
    else:
        out["vsd_keys"] = np.array(list(vsd.keys()))
        out["tsd_keys"] = np.array(list(tsd.keys()))
        out["vsd_shapes"] = np.array([str(tuple(v.shape)) for v in vsd.values()])
        out["tsd_shapes"] = np.array([str(tuple(v.shape)) for v in tsd.values()])
    print(tag, "vis", out["vis_emb"].shape, "txt", out["txt_emb"].shape)
    return out


def run_frame_case(mm, D, H, dims, rows, frames, seed, ragged=False):
    """FrameLAFF config: VisMutiTransformNetPlusFrameFeat (frame-level Attention_1 + 4 video features, batch_norm)."""
    import torch
    vis_dims = {synth.VIS_C3D: dims["c3d"], synth.VIS_TF: dims["tf"], synth.VIS_X3D: dims["x3d"],
                synth.VIS_IRCSN: dims["ircsn"], synth.VIS_FRAME: dims["clip"]}
    cfg = make_config("frame", D, H, vis_dims, dims)
    net = mm.VisMutiTransformNetPlusFrameFeat(cfg)
    sd = load_synth_state(net, seed)
    vis_in = {name: synth.feature(seed, "vis/" + name, rows, d, "relu") for name, d in vis_dims.items() if name != synth.VIS_FRAME}
    fr = synth.feature(seed, "frames", rows * frames, dims["clip"]).reshape(rows, frames, dims["clip"])
    mask = np.ones((rows, frames), dtype=np.float32)
    if ragged:  # zero-padded tails as data_provider.collate_vision produces (data_provider.py:46-61)
        lens = synth.rng_for(seed, "lens").randint(1, frames + 1, size=rows)
        lens[0] = frames  # the reference slices with mask_tensor[0].sum() on the batch axis (a no-op), keep it maximal
        for i, n in enumerate(lens):
            fr[i, n:] = 0
            mask[i, n:] = 0
    with torch.no_grad():
        emb = net({k: torch.from_numpy(v) for k, v in vis_in.items()},
                  {"mask_tensor": torch.from_numpy(mask), synth.VIS_FRAME: torch.from_numpy(fr.copy())})
        # the frame-level stage alone (model/model.py:2167-2173), per video
        fa = net.frame_attention[synth.VIS_FRAME]
        frame_emb = torch.cat([fa(torch.from_numpy(fr[i:i + 1])) for i in range(rows)], 0)
    out = {"meta": np.array([D, H, rows, frames, seed, int(ragged)]), "emb": emb.numpy(), "frame_emb": frame_emb.numpy(),
           "frames": fr, "mask": mask, "names": np.array(list(vis_dims.keys())), "dims": np.array(list(vis_dims.values()))}
    for k, v in vis_in.items():
        out["vin/" + k] = v
    for k, v in sd.items():
        out["sd/" + k] = v
    print("frame", emb.shape, "ragged", ragged)
    return out


def run_attention_variants(ratt, seed):
    """Multi_head_MyApply_Attention / Attention_1 for every (with_ave, mul) on a given stacked [B, L, D] tensor."""
    import torch
    B, L, H, dh = 5, 4, 8, 32
    out = {}
    Y = np.tanh(synth.rng_for(seed, "att/Y").standard_normal((B, L, H * dh))).astype(np.float32)
    out["Y"] = Y
    for with_ave in (0, 1):
        for mul in (0, 1):
            m = ratt.Multi_head_MyApply_Attention(H * dh, H, dh, with_ave=bool(with_ave), mul=bool(mul), split_head=True)
            sd = load_synth_state(m, seed + 10 * with_ave + mul, omega=0.6)
            with torch.no_grad():
                o = m(torch.from_numpy(Y))
                w = torch.stack([m.attention_layer[h].weights for h in range(H)], 1)
            tag = "ave%d_mul%d" % (with_ave, mul)
            out[tag + "/out"], out[tag + "/att"] = o.numpy(), w.numpy()
            for k, v in sd.items():
                out[tag + "/sd/" + k] = v
    return out


def run_sim_eval(mm, rloss, reval, seed):
    """get_txt2vis_matrix, argsort ranks, eval_qry2retro, eval (label matrix), numpy cosine_sim, with planted ties."""
    import torch
    out = {}
    Q = V = 64
    H, dh = 8, 32
    q, g, gt = synth.retrieval_embeddings(seed, Q, V, H, dh, sigma=1.5)
    gt = np.arange(Q)  # eval_qry2retro(n_qry=1) assumes the diagonal
    q = synth.unit_heads(g + 1.5 * synth.unit_heads(synth.rng_for(seed, "n").standard_normal(g.shape).astype(np.float32), H), H)
    g[7] = g[3]       # exact ties: videos 3 and 7 identical
    g[40] = g[41]
    qb, gb = synth.bf16_round(q), synth.bf16_round(g)  # what the tensor-core path stores
    model = mm.W2VVPP(None)
    with torch.no_grad():
        s = model.get_txt2vis_matrix(torch.from_numpy(qb).view(Q, H, dh), torch.from_numpy(gb).view(V, H, dh)).numpy()
        s_raw = model.get_txt2vis_matrix(torch.from_numpy(q).view(Q, H, dh), torch.from_numpy(g).view(V, H, dh)).numpy()
    inds = np.argsort(s, axis=1)
    ranks = np.array([np.where(inds[i][::-1] == i)[0][0] for i in range(Q)])
    out.update(q=q, g=g, q_bf16=qb, g_bf16=gb, scores_bf16=s, scores_fp32=s_raw, argsort=inds, rank0=ranks)
    out["eval_qry2retro"] = np.array(reval.eval_qry2retro(s, n_qry=1), dtype=np.float64)
    label = np.zeros_like(s)
    for i in range(Q):  # predictor.py:239-244
        label[i][np.where(inds[i][::-1] == i)[0]] = 1
    out["eval_label"] = np.array(reval.eval(label), dtype=np.float64)
    out["np_cosine"] = reval.cosine_sim.__wrapped__(q, g) if hasattr(reval.cosine_sim, "__wrapped__") else reval.cosine_sim(q, g)
    x = synth.rng_for(seed, "l2").standard_normal((9, 40)).astype(np.float32)
    x[4] = 0
    out["l2_in"] = x
    out["l2_torch"] = rloss.l2norm(torch.from_numpy(x)).numpy()
    out["l2_torch_eps0"] = rloss.l2norm(torch.from_numpy(x), eps=0).numpy()
    out["l2_numpy"] = reval.l2norm(x)
    # metric edge cases: odd/even Q, all rank 0, big ranks
    for name, rk in (("odd", [0, 3, 1, 10, 2]), ("even", [0, 3, 1, 10, 2, 7]), ("zeros", [0, 0, 0, 0]), ("big", [999999, 5, 123456, 9])):
        rk = np.array(rk)
        n = len(rk)
        Vn = int(rk.max()) + 2
        sim = np.zeros((n, n))  # build a sim matrix whose diagonal has the requested rank is costly for big V: use eval() path
        lab = np.zeros((n, Vn))
        lab[np.arange(n), rk] = 1
        out["metrics_%s/rank0" % name] = rk
        out["metrics_%s/eval" % name] = np.array(reval.eval(lab), dtype=np.float64)
    return out


def run_loss(rloss, seed):
    import torch
    out = {}
    B, H, dh = 16, 8, 32
    r = synth.rng_for(seed, "loss")
    vis = r.standard_normal((B, H, dh)).astype(np.float32)
    txt = (vis + 0.8 * r.standard_normal((B, H, dh))).astype(np.float32)
    out["txt"], out["vis"] = txt, vis
    for mv in (1, 0):
        for direction in ("t2i", "i2t", "bidir"):
            for style in ("sum", "mean"):
                crit = rloss.MarginRankingLoss(margin=0.2, measure="cosine", max_violation=bool(mv), cost_style=style,
                                               direction=direction)
                t = torch.from_numpy(txt).clone().requires_grad_(True)
                v = torch.from_numpy(vis).clone().requires_grad_(True)
                total = 0
                for h in range(H):  # model/model.py:857-858
                    total = total + crit(t[:, h, :], v[:, h, :])
                total.backward()
                tag = "mv%d_%s_%s" % (mv, direction, style)
                out[tag + "/loss"] = total.detach().numpy()
                out[tag + "/d_txt"], out[tag + "/d_vis"] = t.grad.numpy(), v.grad.numpy()
    sc = (0.2 * r.standard_normal((B, B))).astype(np.float32)
    out["score"] = sc
    for mv in (1, 0):
        for direction in ("t2i", "bidir"):
            crit = rloss.MarginRankingLossWithScore(margin=0.2, max_violation=bool(mv), cost_style="sum", direction=direction)
            s = torch.from_numpy(sc).clone().requires_grad_(True)
            l = crit(s)
            l.backward()
            tag = "score_mv%d_%s" % (mv, direction)
            out[tag + "/loss"], out[tag + "/d_score"] = l.detach().numpy(), s.grad.numpy()
    return out


SMALL = dict(clip=32, gru=40, bow=56, w2v=20, x3d=40, ircsn=48, tf=24, c3d=40)


def main():
    import torch
    torch.set_num_threads(8)
    mm, rloss, reval, ratt = import_reference()
    save = lambda name, d: np.savez_compressed(os.path.join(HERE, name), **d)
    # small-dimension cases: inputs, parameters and outputs all stored
    save("fusion_small.npz", run_fusion_case(mm, "small", 256, 8, SMALL, rows=6, seed=11))
    save("fusion_small_ave_mul.npz", run_fusion_case(mm, "small+ave+mul", 256, 8, SMALL, rows=6, seed=12, with_ave=True, mul=True, omega=0.6))
    save("fusion_small_bf16in.npz", run_fusion_case(mm, "small bf16 operands", 256, 8, SMALL, rows=6, seed=13, bf16_inputs=True))
    save("frame_small.npz", run_frame_case(mm, 256, 8, SMALL, rows=5, frames=7, seed=21))
    save("frame_small_ragged.npz", run_frame_case(mm, 256, 8, SMALL, rows=5, frames=9, seed=22, ragged=True))
    save("attention_variants.npz", run_attention_variants(ratt, 31))
    save("sim_eval.npz", run_sim_eval(mm, rloss, reval, 41))
    save("loss.npz", run_loss(rloss, 51))
    # full-dimension case (D=4096, H=8, real feature dims): only outputs stored, inputs/params regenerate from seeds
    full = run_fusion_case(mm, "full", 4096, 8, synth.DIMS, rows=4, seed=61, bf16_inputs=True, store_params=False)
    save("fusion_full_bf16in.npz", full)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
