"""Plain config objects carrying the fields of the reference's config classes that the hot path reads.

The drop-in modules accept the reference's own config objects (configs/laff.py, configs/FrameLaff_..., after
trainer.prepare_config has filled vis_fc_layers[0] / txt_fc_layers / t2v_bow / t2v_w2v); these factories build
equivalent objects without the reference tree, with the values the two shipped LAFF scripts resolve to.
"""
from __future__ import annotations

import types
from typing import Mapping, Optional

from . import synth


class LaffConfig:
    # defaults: configs/base_config.py
    model_name = "LAFF"
    dropout = 0.2                      # base_config.py:79
    activation = "tanh"                # base_config.py:82
    batch_norm = False                 # base_config.py:72
    loss = "mrl"                       # base_config.py:84
    margin = 0.2                       # base_config.py:85
    direction = "t2i"                  # base_config.py:86
    max_violation = True               # base_config.py:88
    cost_style = "sum"                 # base_config.py:90
    measure = "cosine"                 # base_config.py:92
    optimizer = "rmsprop"              # base_config.py:95
    lr = 0.0001                        # base_config.py:97
    lr_decay_rate = 0.99               # base_config.py:98
    grad_clip = 2                      # base_config.py:100
    negative = False                   # base_config.py:245
    float16 = False
    multi_space = True                 # base_config.py:167
    attention_l2norm = False           # base_config.py:125
    vis_expert_embedding = {"expert": False, "l2norm": False}
    txt_expert_embedding = {"expert": False, "l2norm": False}
    vis_feat_add_concat = False
    vis_attention_global_decay_rate = 0.8
    txt_attention_global_decay_rate = 0.8
    frame_feat_with_video_feat = True
    vis_frame_addFC = False
    vid_frame_feats = ()


def _text_encoding():
    return {"bow_encoding": {"name": "bow_nsw"}, "w2v_encoding": {"name": "w2v_nsw"},
            "rnn_encoding": {"name": "gru_mean"}, "bert_encoding": {"name": "noBert"},
            "CLIP_encoding": {"name": "ViT-B/32"}, "NetVLAD_encoding": {"name": "noNetVLAD"}}


def laff_config(D: int = 4096, heads: int = 8, dims: Optional[Mapping[str, int]] = None, with_ave: bool = False,
                mul: bool = False) -> LaffConfig:
    """configs/laff.py with adjust_parm('0_12_0_12_<ave>_<mul>_1'): video = clip-ft (no transform) + TimeSformer + X3D
    + irCSN, text = gru + bow + w2v + CLIP (no transform), Multi_head_MyApply_Attention on both sides."""
    d = dict(synth.DIMS if dims is None else dims)
    c = LaffConfig()
    c.model_name = "LAFF"
    c.vis_fc_layers = [{synth.VIS_CLIP_FT: d["clip"], synth.VIS_TF: d["tf"], synth.VIS_X3D: d["x3d"],
                        synth.VIS_IRCSN: d["ircsn"]}, D]
    c.txt_fc_layers = [0, D]
    c.text_encoding = _text_encoding()
    c.clip_opt = {"size": d["clip"], "transform_batch_norm": True, "transform_dropout": 0.0,
                  "transform_activation": "tanh", "frozen": True}          # configs/laff.py:35-38
    c.rnn_size = d["gru"]
    c.t2v_bow = types.SimpleNamespace(ndims=d["bow"])
    c.t2v_w2v = types.SimpleNamespace(ndims=d["w2v"])
    c.multi_head_attention = {"dropout": 0.0, "heads": heads, "embed_dim_qkv": D // heads}   # configs/laff.py:43-46
    c.attention_param_each_head = {"with_ave": with_ave, "mul": mul, "split_head": True}     # configs/laff.py:86-88
    c.vis_attention = c.txt_attention = "Multi_head_MyApply_Attention"                       # attention_types[12]
    c.vis_no_transform = [synth.VIS_CLIP_FT]                                                 # configs/laff.py:49
    c.txt_no_transform = ["CLIP_encoder"]                                                    # configs/laff.py:50
    c.vid_feats = list(c.vis_fc_layers[0].keys())                                            # configs/laff.py:28-33
    return c


def frame_laff_config(D: int = 4096, heads: int = 8, dims: Optional[Mapping[str, int]] = None) -> LaffConfig:
    """configs/FrameLaff_NoFrameFc_StrongCLIP_adjust.py with adjust_parm('0_7_1_12_0_12_0') (LAFF-ml): frame feature
    pooled by Attention_1(with_ave=False, mul=False), video-level C3D + TimeSformer + X3D + irCSN, batch_norm=True."""
    d = dict(synth.DIMS if dims is None else dims)
    c = laff_config(D, heads, d)
    c.model_name = "FrameLAFF"
    c.batch_norm = True                                                     # FrameLaff...:10
    c.float16 = True                                                        # FrameLaff...:33 (AMP in training only)
    c.vis_fc_layers = [{synth.VIS_C3D: d["c3d"], synth.VIS_TF: d["tf"], synth.VIS_X3D: d["x3d"],
                        synth.VIS_IRCSN: d["ircsn"], synth.VIS_FRAME: d["clip"]}, D]
    c.vid_frame_feats = [synth.VIS_FRAME]                                   # FrameLaff...:87
    c.vis_no_transform = [synth.VIS_FRAME]                                  # FrameLaff...:88
    c.vis_frame_attention = "attention_noAveNoAverageMul"                   # attention_types[7]
    c.vis_frame_addFC = False                                               # FrameLaff...:58
    c.frame_feat_with_video_feat = True                                     # FrameLaff...:53
    c.attention_param_each_head = {"with_ave": False, "mul": False, "split_head": True}
    c.vis_attention_global_decay_rate = 0.0
    c.txt_attention_global_decay_rate = 0.0
    return c
