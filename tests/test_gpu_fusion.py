"""GPU parity of the fusion path against (a) golden vectors from the unmodified reference and
(b) the CPU oracle, through the drop-in modules (which call the C ABI).

Tolerances (stated):
  T1  same stage inputs (bf16-representable operands): fused embedding abs err <= 2e-5 per component at small dims,
      <= 2e-6 at D=4096 (components ~0.04); reference fp32 vs tensor-core fp32 accumulation only.
  T2  fp32 inputs through bf16 operands vs the all-fp32 reference: bound 6e-3 at d_h = 32 (2^-9 relative per operand);
      with the 3-term split ('bf16x3') the products are near-fp32 and T1 applies.
"""
import numpy as np
import pytest
import torch

from helpers import cuda, load_numpy_state, max_abs, sd_from_npz, small_dims
from laff_b200 import config as cfg
from laff_b200 import loss as L
from laff_b200 import model as M
from laff_b200 import ops, synth
from oracle import laff_oracle as O
from test_oracle_golden import regen_full

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _default_precision():
    L.set_precision("bf16")
    yield
    L.set_precision("bf16")


def build_laff_nets(d, D, H, dims, with_ave=False, mul=False):
    c = cfg.laff_config(D, H, dims, with_ave, mul)
    vis = M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0]).cuda()
    txt = M.MultiScaleTxtEncoderAttention(c).cuda()
    return vis, txt


@pytest.mark.parametrize("fname,precision,tol", [
    ("fusion_small_bf16in.npz", "bf16", 2e-5),      # T1: operands exactly representable
    ("fusion_small.npz", "bf16x3", 2e-5),            # near-fp32 products on unrounded fp32 inputs
    ("fusion_small.npz", "bf16", 6e-3),              # T2 (components ~0.18 at d_h = 32)
    ("fusion_small.npz", "fp16", 8e-4),              # T2 with fp16 operands (8x finer rounding)
    ("fusion_small_ave_mul.npz", "bf16x3", 2e-5),    # with_ave + mul variant, omega = 0.6
])
def test_fusion_small_vs_reference(golden, fname, precision, tol):
    d = golden(fname)
    D, H, rows, seed, with_ave, mul, _ = [int(x) for x in d["meta"]]
    vis, txt = build_laff_nets(d, D, H, small_dims(d), bool(with_ave), bool(mul))
    load_numpy_state(vis, sd_from_npz(d, "vsd/"))
    load_numpy_state(txt, sd_from_npz(d, "tsd/"))
    L.set_precision(precision)
    names = [str(n) for n in d["vis_names"]]
    v = vis({n: torch.from_numpy(d["vin/" + n]) for n in names})            # CPU tensors in, like the reference's loaders
    t = txt({k: torch.from_numpy(d["tin/" + k]) for k in ("gru", "bow", "w2v", "clip")})
    assert v.shape == (rows, H, D // H) and v.is_cuda and v.dtype == torch.float32
    assert max_abs(v, d["vis_emb"]) <= tol
    assert max_abs(t, d["txt_emb"]) <= tol
    if precision == "bf16x3" or "bf16in" in fname:
        # attention weights the reference keeps in Attention_1.weights
        va = torch.stack([vis.attention_layer.attention_layer[h].weights for h in range(H)], 1)
        assert max_abs(va, d["vis_att"]) <= 5e-5
    np.testing.assert_allclose(v.norm(dim=2).cpu().numpy(), 1.0, atol=1e-5)


def test_fusion_full_dims_vs_reference_and_oracle(golden):
    """D = 4096, H = 8, real feature dims (clip 512, tf 768, x3d/ircsn 2048, gru 1024, bow 3981, w2v 500)."""
    d = golden("fusion_full_bf16in.npz")
    H, vis_in, txt_in, vsd, tsd = regen_full(d)
    vis, txt = build_laff_nets(d, 4096, 8, synth.DIMS)
    load_numpy_state(vis, vsd)
    load_numpy_state(txt, tsd)
    v = vis({k: torch.from_numpy(x) for k, x in vis_in.items()})
    t = txt({k: torch.from_numpy(x) for k, x in txt_in.items()})
    assert max_abs(v, d["vis_emb"]) <= 2e-6 and max_abs(t, d["txt_emb"]) <= 2e-6
    ov, _ = O.vis_net_forward(vis_in, vsd, [synth.VIS_CLIP_FT], H)
    ot, _ = O.txt_net_forward(txt_in, tsd, ["CLIP_encoder"], H)
    assert max_abs(v, ov) <= 2e-6 and max_abs(t, ot) <= 2e-6
    # 16-bit copy written by the same kernel = rounding of the fp32 output
    v32, v16 = vis.encode({k: torch.from_numpy(x) for k, x in vis_in.items()}, out16_dtype=torch.bfloat16)
    assert torch.equal(v16, v32.to(torch.bfloat16))


@pytest.mark.parametrize("fname", ["frame_small.npz", "frame_small_ragged.npz"])
def test_frame_laff_vs_reference(golden, fname):
    d = golden(fname)
    D, H = int(d["meta"][0]), int(d["meta"][1])
    names = [str(n) for n in d["names"]]
    dims = dict(zip(names, [int(x) for x in d["dims"]]))
    sm = {"clip": dims[synth.VIS_FRAME], "c3d": dims[synth.VIS_C3D], "tf": dims[synth.VIS_TF], "x3d": dims[synth.VIS_X3D],
          "ircsn": dims[synth.VIS_IRCSN], "gru": 40, "bow": 56, "w2v": 20}
    c = cfg.frame_laff_config(D, H, sm)
    net = M.VisMutiTransformNetPlusFrameFeat(c).cuda()
    load_numpy_state(net, sd_from_npz(d, "sd/"))
    L.set_precision("bf16x3")
    vis_in = {n: torch.from_numpy(d["vin/" + n]) for n in names if n != synth.VIS_FRAME}
    frames = {"mask_tensor": torch.from_numpy(d["mask"]), synth.VIS_FRAME: torch.from_numpy(d["frames"])}
    fe = net.frame_pool(synth.VIS_FRAME, cuda(d["frames"]))
    assert max_abs(fe, d["frame_emb"]) <= 2e-6                      # F7 alone (HBM-bound kernel, fp32)
    emb = net(vis_in, frames)
    assert max_abs(emb, d["emb"]) <= 2e-5


def test_frame_pool_is_pad_invariant_and_matches_oracle():
    """32 frames x 512 (config C4) + ragged zero padding: padded frames take part in the softmax exactly like the
    reference (model/model.py:2167-2173) and cancel under the L2 norm."""
    r = synth.rng_for(3, "fp")
    B, F, dim = 300, 32, 512
    fr = r.standard_normal((B, F, dim)).astype(np.float32)
    w = (r.standard_normal(dim) / np.sqrt(dim)).astype(np.float32)
    lens = r.randint(1, F + 1, size=B)
    for i, n in enumerate(lens):
        fr[i, n:] = 0
    out = ops.frame_pool(cuda(fr), cuda(w), 0.03)
    ref, _ = O.attention_1(fr, w, np.float32(0.03), with_ave=False, mul=False)
    assert max_abs(out, ref) <= 2e-6
    unpadded = np.stack([O.attention_1(fr[i:i + 1, :lens[i]], w, np.float32(0.03), False, False)[0][0] for i in range(B)])
    assert max_abs(out, unpadded) <= 2e-6


def test_attention_variants_vs_reference(golden):
    d = golden("attention_variants.npz")
    Y = cuda(d["Y"])
    for with_ave in (0, 1):
        for mul in (0, 1):
            tag = "ave%d_mul%d" % (with_ave, mul)
            m = M.Multi_head_MyApply_Attention(256, 8, 32, with_ave=bool(with_ave), mul=bool(mul), split_head=True).cuda()
            load_numpy_state(m, sd_from_npz(d, tag + "/sd/"))
            out = m(Y)
            assert max_abs(out, d[tag + "/out"]) <= 2e-6, tag
            att = torch.stack([m.attention_layer[h].weights for h in range(8)], 1)
            assert max_abs(att, d[tag + "/att"]) <= 2e-6, tag
    # the single-head block on its own (frame attention type)
    a1 = M.Attention_1(32, with_ave=False, mul=False).cuda()
    sd = sd_from_npz(d, "ave0_mul0/sd/")
    a1.load_state_dict({k[len("attention_layer.0."):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith("attention_layer.0.")})
    o = a1(Y[:, :, :32].contiguous())
    assert max_abs(o, d["ave0_mul0/out"][:, 0, :]) <= 2e-6


def test_state_dict_keys_match_reference(golden):
    d = golden("fusion_full_bf16in.npz")
    vis, txt = build_laff_nets(d, 4096, 8, synth.DIMS)
    assert list(vis.state_dict().keys()) == [str(k) for k in d["vsd_keys"]]
    assert sorted(txt.state_dict().keys()) == sorted(str(k) for k in d["tsd_keys"])
    for k, s in zip(d["vsd_keys"], d["vsd_shapes"]):
        assert tuple(vis.state_dict()[str(k)].shape) == eval(str(s))
    f = golden("frame_small.npz")
    D, H = int(f["meta"][0]), int(f["meta"][1])
    names = [str(n) for n in f["names"]]
    dims = dict(zip(names, [int(x) for x in f["dims"]]))
    sm = {"clip": dims[synth.VIS_FRAME], "c3d": dims[synth.VIS_C3D], "tf": dims[synth.VIS_TF], "x3d": dims[synth.VIS_X3D],
          "ircsn": dims[synth.VIS_IRCSN], "gru": 40, "bow": 56, "w2v": 20}
    net = M.VisMutiTransformNetPlusFrameFeat(cfg.frame_laff_config(D, H, sm))
    assert list(net.state_dict().keys()) == list(sd_from_npz(f, "sd/").keys())


def test_projection_edge_shapes_vs_oracle():
    """TransformNet alone: K not a multiple of 64 (w2v 500, bow 3981), rows not a multiple of 128, every activation,
    with and without BN; both CTA-group modes of the GEMM engine."""
    r = synth.rng_for(9, "proj")
    before = ops.get_tuning()
    try:
        for cg in (1, 2):
            ops.set_tuning(cta_group=cg)
            for rows, K, act, bn in ((1, 500, "tanh", False), (130, 3981, "tanh", True), (257, 24, "relu", True),
                                     (300, 768, "sigmoid", False), (129, 1024, None, False)):
                tn = M.TransformNet((K, 4096), None, 0.0, bn, act).cuda().eval()
                with torch.no_grad():
                    tn.fc1.weight.copy_(cuda(synth.bf16_round(r.uniform(-0.05, 0.05, (4096, K)).astype(np.float32))))
                    tn.fc1.bias.copy_(cuda(r.standard_normal(4096).astype(np.float32) * 0.1))
                    if bn:
                        tn.bn1.running_mean.copy_(cuda(r.standard_normal(4096).astype(np.float32) * 0.3))
                        tn.bn1.running_var.copy_(cuda(r.uniform(0.5, 2, 4096).astype(np.float32)))
                        tn.bn1.weight.copy_(cuda(r.uniform(0.5, 1.5, 4096).astype(np.float32)))
                x = synth.bf16_round(r.standard_normal((rows, K)).astype(np.float32))
                y = tn(torch.from_numpy(x))
                sd = {k: v.detach().cpu().numpy() for k, v in tn.state_dict().items()}
                ref = O.transform_net(x.astype(np.float64), {k: v.astype(np.float64) if v.dtype.kind == "f" else v for k, v in sd.items()}, "", act)
                assert y.shape == (rows, 4096)
                assert max_abs(y, ref) <= 3e-5, (cg, rows, K, act, bn)
    finally:
        ops.set_tuning(*before)


def test_operand_preparation_kernels():
    r = synth.rng_for(10, "prep")
    x = r.standard_normal((37, 500)).astype(np.float32)
    c = ops.cast_pad_16(cuda(x), torch.bfloat16)
    assert c.shape == (37, 504) and torch.equal(c[:, :500].float().cpu(), torch.from_numpy(synth.bf16_round(x)))
    assert bool((c[:, 500:] == 0).all())
    xl, xr = ops.split3_16(cuda(x), 0), ops.split3_16(cuda(x), 1)
    approx = xl.double() @ xr.double().T
    exact = torch.from_numpy(x).double() @ torch.from_numpy(x).double().T
    assert float((approx.cpu() - exact).abs().max() / exact.abs().max()) < 2e-5
    e = r.standard_normal((9, 8 * 64)).astype(np.float32)
    e[4] = 0
    n = ops.l2norm_quantize(cuda(e), 8, torch.float32)
    ref = O.l2norm(e.reshape(9 * 8, 64)).reshape(9, 512)
    assert max_abs(n, ref) <= 2e-7 and bool((n[4] == 0).all())
    nb = ops.l2norm_quantize(cuda(e), 8, torch.bfloat16)
    assert torch.equal(nb.float().cpu(), torch.from_numpy(synth.bf16_round(n.cpu().numpy())))
    assert max_abs(L.l2norm(cuda(e[:, :64])), O.l2norm(e[:, :64])) <= 2e-7


def test_train_mode_is_refused_not_faked():
    tn = M.TransformNet((64, 256), None, 0.2, True, "tanh").cuda().train()
    with pytest.raises(NotImplementedError):
        tn(torch.randn(4, 64))
    with pytest.raises(NotImplementedError):
        M.get_attention_layer("muti_head_attention_official", 256, 4, cfg.laff_config(256, 8))


@pytest.fixture
def fuse_variant(request):
    ops.set_fuse_variant(request.param)
    yield request.param
    ops.set_fuse_variant(0)


@pytest.mark.parametrize("fuse_variant", [1, 2], indirect=True)
@pytest.mark.parametrize("kind", ["laff_txt", "laff_vis", "frame_vis"])
def test_single_kernel_fusion_vs_two_kernel_path_and_oracle(kind, fuse_variant):
    """laff_fuse_forward (all projections + pooling in one kernel; both the cta_group::1 / 2-CTA-cluster and the
    cta_group::2 / 4-CTA-cluster variants) against the two-kernel path and the oracle, at the real dimensions, for
    ragged row counts, with and without BatchNorm (LAFF-ml has BN on every FC feature)."""
    from laff_b200 import _capi
    assert ops.get_fuse_variant() == fuse_variant
    r = synth.rng_for(21, kind)
    if kind == "frame_vis":
        c = cfg.frame_laff_config(4096, 8, synth.DIMS)
        net = M.VisMutiTransformNetPlusFrameFeat(c)
    else:
        c = cfg.laff_config(4096, 8, synth.DIMS)
        net = M.MultiScaleTxtEncoderAttention(c) if kind == "laff_txt" else M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0])
    sd = {k: synth.bf16_round(synth.param(3, k, tuple(v.shape))) if k.endswith("fc1.weight") else synth.param(3, k, tuple(v.shape))
          for k, v in net.state_dict().items()}
    load_numpy_state(net, sd)
    net = net.cuda().eval()
    for rows in (1, 127, 129, 300, 700):
        if kind == "laff_txt":
            feats = {"gru": synth.bf16_round(r.standard_normal((rows, 1024)).astype(np.float32)),
                     "bow": r.randint(0, 3, (rows, 3981)).astype(np.float32),
                     "w2v": synth.bf16_round(r.standard_normal((rows, 500)).astype(np.float32)),
                     "clip": r.standard_normal((rows, 512)).astype(np.float32)}
            ref, _ = O.txt_net_forward(feats, sd, ["CLIP_encoder"], 8)
            run = lambda: net.encode({k: torch.from_numpy(v) for k, v in feats.items()}, out16_dtype=torch.bfloat16)
        elif kind == "laff_vis":
            feats = {k: (r.standard_normal((rows, d)).astype(np.float32) if k == synth.VIS_CLIP_FT else
                         synth.bf16_round(np.maximum(r.standard_normal((rows, d)), 0).astype(np.float32)))
                     for k, d in c.vis_fc_layers[0].items()}
            ref, _ = O.vis_net_forward(feats, sd, [synth.VIS_CLIP_FT], 8)
            run = lambda: net.encode({k: torch.from_numpy(v) for k, v in feats.items()}, out16_dtype=torch.bfloat16)
        else:
            feats = {k: synth.bf16_round(np.maximum(r.standard_normal((rows, d)), 0).astype(np.float32))
                     for k, d in c.vis_fc_layers[0].items() if k != synth.VIS_FRAME}
            frames = r.standard_normal((rows, 6, 512)).astype(np.float32)
            ref, _ = O.frame_vis_net_forward(feats, frames, synth.VIS_FRAME, sd, [synth.VIS_FRAME], 8)
            fd = {"mask_tensor": torch.ones(rows, 6), synth.VIS_FRAME: torch.from_numpy(frames)}
            run = lambda: net.encode({k: torch.from_numpy(v) for k, v in feats.items()}, fd, out16_dtype=torch.bfloat16)
        lib = _capi.lib()
        M.set_single_kernel_fusion(True)
        run()                                                     # first call casts / folds the weights (cached)
        lib.laff_launch_count(1)
        a32, a16 = run()
        n_single = lib.laff_launch_count(1)
        M.set_single_kernel_fusion(False)
        b32, b16 = run()
        n_two = lib.laff_launch_count(1)
        M.set_single_kernel_fusion(True)
        assert n_single < n_two                                   # the single-kernel path really is the one that ran
        assert max_abs(a32, ref) <= 2e-6 and max_abs(b32, ref) <= 2e-6, (kind, rows)
        assert max_abs(a32, b32) <= 5e-7
        assert torch.equal(a16, a32.to(torch.bfloat16))
        np.testing.assert_allclose(a32.norm(dim=2).cpu().numpy(), 1.0, atol=1e-6)


def test_fuse_forward_rejects_unsupported_shapes():
    from laff_b200 import LaffError
    x16 = torch.zeros(4, 64, dtype=torch.bfloat16, device="cuda")
    w16 = torch.zeros(256, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(LaffError, match="head_dim must be 512"):
        ops.fuse_forward([{"x16": x16, "w16": w16, "bias": None, "activation": "tanh"}], [], torch.zeros(8, 32, device="cuda"),
                         torch.zeros(8, device="cuda"), 8, 32)
