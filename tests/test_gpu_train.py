"""Training step on the GPU: model(train_data) — train-mode forward, loss, explicit
backward, gradient clipping and RMSprop / Adam through the C ABI — against three real steps of the unmodified
reference model (tests/golden/train_*.npz) and the oracle.

Bars (3-term bf16-split GEMM operands, fp32 everywhere else): loss rel err <= 2e-5; first-step gradients within 1e-4
of the largest entry of each tensor; parameters after each step within the same bounds the oracle meets on CPU
(the zero-gradient logit biases excepted, see tests/test_train_cpu.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import load_numpy_state
from laff_b200 import config as cfg
from laff_b200 import model as M
from laff_b200 import ops, synth
from laff_b200.train import DeviceOptimizer
from oracle import laff_oracle as O
from test_train_cpu import load_case, step_inputs

pytestmark = pytest.mark.gpu
SMALL = dict(clip=32, gru=40, bow=56, w2v=20, x3d=40, ircsn=48, tf=24, c3d=40)


def build_model(g, sd, H, D):
    from test_train_cpu import variant
    with_ave, mul, loss_kind = variant(g)
    c = cfg.laff_config(D, H, SMALL, with_ave=with_ave, mul=mul)
    c.loss = loss_kind
    c.dropout = 0.0
    c.batch_norm = bool(int(g["meta"][5]))
    c.optimizer, c.lr, c.grad_clip = str(g["optimizer"]), float(g["lr"]), float(g["grad_clip"])
    model = M.get_model("LAFF", torch.device("cuda"), c)
    load_numpy_state(model, sd)
    return model


def train_data(vis_in, txt_in):
    return {"vis_feats": {k: torch.from_numpy(v) for k, v in vis_in.items()}, "captions": {k: torch.from_numpy(v) for k, v in txt_in.items()},
            "captions_task2": None, "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}


@pytest.mark.parametrize("tag", ["rmsprop", "adam", "rmsprop_bn", "rmsprop_ave_mul", "adam_dsl"])
def test_train_steps_match_reference(tag):
    g, sd, H, steps = load_case(tag)
    D = int(g["meta"][1])
    model = build_model(g, sd, H, D).train()
    lr = float(g["lr"])
    for s in range(steps):
        vis_in, txt_in = step_inputs(g, s)
        items = model(train_data(vis_in, txt_in), epoch=0)
        loss = float(items["triplet_loss"])
        assert abs(loss - g["losses"][s]) <= 2e-5 * abs(g["losses"][s]), (tag, s, loss, g["losses"][s])
        if s == 0:
            grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
            ref_keys = [k[6:] for k in g.files if k.startswith("grad0/")]
            assert sorted(ref_keys) == sorted(grads.keys())
            for k in ref_keys:
                ref = g["grad0/" + k]
                got = grads[k].cpu().numpy().reshape(ref.shape)          # clipped in place, like clip_grad_norm_
                assert np.abs(got - ref).max() <= 1e-4 * max(1e-3, np.abs(ref).max()), (k, np.abs(got - ref).max())
        cur = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        for k, v in cur.items():
            ref = g["sd%d/%s" % (s + 1, k)]
            err = np.abs(v.reshape(ref.shape).astype(np.float64) - ref)
            # RMSprop / Adam normalise the gradient: an element whose gradient is at the level of the fp32 gradient
            # noise (|g| ~ 1e-7, e.g. weights fed by ReLU-zero inputs, or the shift-invariant logit bias) moves by up
            # to lr / sqrt(1 - alpha) in a direction the noise decides — in the reference as much as here.  So: every
            # element within that bound, and the elements with a well-resolved first-step gradient tight.
            assert err.max() <= 11 * lr * (s + 1) * max(1.0, np.abs(ref).max()), (k, s, err.max())
            if s == 0 and "grad0/" + k in g.files and not k.endswith("embedding_common.0.bias"):
                g0 = np.abs(g["grad0/" + k]).reshape(ref.shape)
                solid = g0 > 1e-2 * g0.max()
                assert solid.any() and err[solid].max() <= 5e-5 * max(1.0, np.abs(ref).max()), (k, err[solid].max())
            if not k.endswith("embedding_common.0.bias"):
                assert np.mean(err <= 3e-4 * (s + 1)) >= 0.97, (k, s, np.mean(err <= 3e-4 * (s + 1)))
    assert int(model.iters) == steps


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_cuda_graph_replay_equals_eager_steps(dropout):
    """From the 4th step on the whole step is one replayed CUDA graph (device-side step counter, learning rate and
    dropout seed): parameters and losses must equal those of the same steps issued eagerly, bit for bit."""
    g, sd, H, steps = load_case("adam")
    D = int(g["meta"][1])
    runs = []
    for use_graph in (False, True):
        model = build_model(g, sd, H, D).train()
        model.opt.dropout = dropout
        for m in model.modules():
            if isinstance(m, M.TransformNet) and m.fc1 is not None:
                m.dropout_p = dropout if dropout > 0 else None
        model.use_cuda_graph = use_graph
        losses = []
        for s in range(8):
            vis_in, txt_in = step_inputs(g, s % steps)
            if s == 6:
                model.optimizer.param_groups[0]["lr"] *= 0.5          # an lr scheduler acting between steps
            losses.append(float(model(train_data(vis_in, txt_in))["triplet_loss"]))
        assert (model._graph is not None) == use_graph
        runs.append((losses, {k: v.detach().clone() for k, v in model.state_dict().items()}))
    assert runs[0][0] == runs[1][0]
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
    key = "vis_net.VisMutiTransformNet.%s.bn1.num_batches_tracked" % synth.VIS_CLIP_FT
    assert int(runs[1][1][key]) == int(sd[key]) + 8


def test_eval_after_training_uses_the_updated_parameters():
    """The optimizer updates parameters through raw pointers; the eval-mode operand caches must follow."""
    g, sd, H, steps = load_case("rmsprop")
    D = int(g["meta"][1])
    model = build_model(g, sd, H, D)
    vis_in, txt_in = step_inputs(g, 0)
    model.eval()
    before = model.vis_net({k: torch.from_numpy(v) for k, v in vis_in.items()}).clone()
    model.train()
    model(train_data(vis_in, txt_in))
    model.eval()
    after = model.vis_net({k: torch.from_numpy(v) for k, v in vis_in.items()})
    new_sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    ref, _ = O.vis_net_forward(vis_in, {k[len("vis_net."):]: v for k, v in new_sd.items() if k.startswith("vis_net.")},
                               [synth.VIS_CLIP_FT], H)
    assert (after - before).abs().max() > 1e-4
    assert np.abs(after.cpu().numpy() - ref).max() <= 6e-3        # bf16 eval path (T2 tolerance at d_h = 32)
    with pytest.raises(ops.LaffError):
        model(train_data(vis_in, txt_in))                          # a training step needs model.train()


def test_dropout_mask_statistics_and_backward_consistency():
    B, D, p = 256, 512, 0.2
    a = torch.rand(B, D, device="cuda") + 0.5
    y, mask, _, _ = ops.transform_train_forward(a, D, p, seed=1234)
    y2, mask2, _, _ = ops.transform_train_forward(a, D, p, seed=1234)
    y3, mask3, _, _ = ops.transform_train_forward(a, D, p, seed=1235)
    assert torch.equal(mask, mask2) and torch.equal(y, y2) and not torch.equal(mask, mask3)      # counter-based: reproducible
    keep = mask.float().mean().item()
    assert abs(keep - (1 - p)) < 0.01
    assert torch.equal(y, torch.where(mask.bool(), a / (1 - p), torch.zeros_like(a)))
    dy = torch.randn(B, D, device="cuda")
    dz = ops.transform_train_backward(dy, torch.tanh(a), None, mask, p, "tanh", None, None, None)
    ref = torch.where(mask.bool(), dy / (1 - p), torch.zeros_like(dy)) * (1 - torch.tanh(a) ** 2)
    assert (dz - ref).abs().max().item() <= 1e-6


def test_optimizer_matches_torch_and_skips_gradless_parameters():
    torch.manual_seed(0)
    for kind in ("rmsprop", "adam"):
        ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in (5, 3000, 70001)]
        qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        extra = torch.nn.Parameter(torch.randn(7, device="cuda"))          # never gets a gradient
        mine = DeviceOptimizer(ps + [extra], kind=kind, lr=1e-3, eps=1e-4 if kind == "adam" else None, max_grad_norm=2.0)
        ref = torch.optim.RMSprop(qs, lr=1e-3) if kind == "rmsprop" else torch.optim.Adam(qs, lr=1e-3, eps=1e-4)
        e0 = extra.detach().clone()
        for step in range(4):
            gs = [torch.randn_like(p) * (10.0 if step % 2 else 0.01) for p in ps]    # alternately clipped / not clipped
            for p, q, gr in zip(ps, qs, gs):
                if p.grad is None:
                    p.grad = gr.clone()
                else:
                    p.grad.copy_(gr)
                q.grad = gr.clone()
            norm = mine.step()
            tn = torch.nn.utils.clip_grad_norm_(qs, 2.0)
            ref.step()
            assert abs(norm.item() - tn.item()) <= 1e-5 * tn.item()
            for p, q in zip(ps, qs):
                assert (p - q).abs().max().item() <= 2e-6, (kind, step)
                assert (p.grad - q.grad).abs().max().item() <= 1e-6 * max(1.0, q.grad.abs().max().item())
        assert torch.equal(extra, e0)


def test_laff_ml_train_steps_match_reference():
    """'FrameLAFF' (LAFF-ml): frame-level attention backward, gradient through the tiled pooled frame feature, BatchNorm
    on every projected feature — three steps against the unmodified reference model."""
    from test_train_cpu import check_params
    g, sd, H, steps = load_case("frame_rmsprop")
    D = int(g["meta"][1])
    ff = str(g["frame_feat"])
    c = cfg.frame_laff_config(D, H, SMALL)
    c.dropout, c.float16 = 0.0, False
    c.optimizer, c.lr, c.grad_clip = str(g["optimizer"]), float(g["lr"]), float(g["grad_clip"])
    model = M.get_model("FrameLAFF", torch.device("cuda"), c)
    load_numpy_state(model, sd)
    model.train()
    for s in range(steps + 3):                      # three more steps run through the replayed CUDA graph
        vis_in, txt_in = step_inputs(g, s % steps)
        td = train_data(vis_in, txt_in)
        td["vis_frame_feat_dict"] = {"mask_tensor": torch.from_numpy(g["step%d/mask" % (s % steps)]),
                                     ff: torch.from_numpy(g["step%d/frames" % (s % steps)])}
        loss = float(model(td, epoch=0)["triplet_loss"])
        if s >= steps:
            assert np.isfinite(loss)
            continue
        # step 0 is a pure function of the inputs: tight.  Later steps inherit the first RMSprop step, which is sign-SGD
        # (g / sqrt(0.01 g^2)): the ~0.1 % of elements whose gradient lies below the 1e-5 relative noise of the split-bf16
        # products step the other way by 2 * lr / sqrt(1 - alpha), which moves the next losses by ~1e-4 relative.
        assert abs(loss - g["losses"][s]) <= (2e-5 if s == 0 else 5e-4) * abs(g["losses"][s]), (s, loss, g["losses"][s])
        if s == 0:
            grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
            ref_keys = [k[6:] for k in g.files if k.startswith("grad0/")]
            assert sorted(ref_keys) == sorted(grads.keys())
            for k in ref_keys:
                ref = g["grad0/" + k]
                got = grads[k].cpu().numpy().reshape(ref.shape)
                assert np.abs(got - ref).max() <= 1e-4 * max(1e-3, np.abs(ref).max()), (k, np.abs(got - ref).max())
        check_params({k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}, g, s, float(g["lr"]), tight0=5e-5, frac=0.97)
    assert model._graph is not None


@pytest.mark.parametrize("tag", ["frame_amp_rmsprop", "frame_amp_adam"])
def test_laff_ml_float16_branch_follows_reference_scaler(tag):
    """config.float16 = True (the shipped FrameLAFF setting, configs/FrameLaff_...:33): the reference's AMP branch
    (model/model.py:970-989) over 12 steps of the unmodified reference.  Loss-scale trajectory and skipped steps
    identical, losses within fp16 forward rounding, the clipped gradient has norm grad_clip / S and the direction of
    the reference's, parameters follow the reference's trajectory; the last steps run as a replayed CUDA graph."""
    g, sd, H, steps = load_case(tag)
    D = int(g["meta"][1])
    ff = str(g["frame_feat"])
    c = cfg.frame_laff_config(D, H, SMALL)
    c.dropout, c.float16 = 0.0, True
    c.optimizer, c.lr, c.grad_clip = str(g["optimizer"]), float(g["lr"]), float(g["grad_clip"])
    model = M.get_model("FrameLAFF", torch.device("cuda"), c)
    load_numpy_state(model, sd)
    model.train()
    first = int(g["first_executed_step"])
    lr, clip = float(g["lr"]), float(g["grad_clip"])
    S = 65536.0
    for s in range(steps):
        vis_in, txt_in = step_inputs(g, s)
        td = train_data(vis_in, txt_in)
        td["vis_frame_feat_dict"] = {"mask_tensor": torch.from_numpy(g["step%d/mask" % s]), ff: torch.from_numpy(g["step%d/frames" % s])}
        loss = float(model(td, epoch=0)["triplet_loss"])
        assert abs(loss - g["losses"][s]) <= 1e-2 * abs(g["losses"][s]), (s, loss, g["losses"][s])
        assert model.scaler.get_scale() == float(g["scales"][s]), (s, model.scaler.get_scale(), g["scales"][s])
        assert model.scaler.last_step_skipped() == bool(g["skipped"][s]), s
        if s == first:
            grads = {k: p.grad.double().cpu().numpy() for k, p in model.named_parameters() if p.grad is not None}
            n = np.sqrt(sum(float((v ** 2).sum()) for v in grads.values()))
            assert abs(n - clip / S) <= 1e-3 * clip / S, (n, clip / S)
            nref = np.sqrt(sum(float((g[k].astype(np.float64) ** 2).sum()) for k in g.files if k.startswith("grad_first/")))
            for k, got in grads.items():
                ref = g["grad_first/" + k].astype(np.float64).ravel()
                if np.linalg.norm(ref) > 1e-3 * nref:
                    cos = float(ref @ got.ravel() / (np.linalg.norm(ref) * np.linalg.norm(got)))
                    assert cos >= 0.99, (k, cos)
        S = model.scaler.get_scale()
    assert model._graph is not None and model.scaler.skipped_steps() == int(g["skipped"].sum())
    executed = int(steps - g["skipped"].sum())
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            continue
        ref = g["sd%d/%s" % (steps, k)]
        err = np.abs(v.detach().cpu().numpy().reshape(ref.shape).astype(np.float64) - ref)
        assert err.max() <= 11 * lr * executed, (k, err.max())
        moved = np.abs(ref - g["sd0/" + k].reshape(ref.shape))
        if moved.max() > 0 and err.size >= 1024:
            assert np.median(err) <= 0.25 * max(np.median(moved), 1e-7), (k, np.median(err), np.median(moved))


def test_optimizer_rebuild_keeps_state_and_step_count():
    """A re-allocated gradient makes DeviceOptimizer rebuild its descriptor table: the running averages and the
    device-side step count must survive (Adam's bias correction depends on it)."""
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(3000, device="cuda"))
    q = torch.nn.Parameter(p.detach().clone())
    a = DeviceOptimizer([p], kind="adam", lr=1e-2, eps=1e-8)
    b = torch.optim.Adam([q], lr=1e-2, eps=1e-8)
    for it in range(4):
        gr = torch.randn(3000, device="cuda")
        p.grad = gr.clone()          # a fresh tensor every step: the pointers change, the optimizer rebuilds
        q.grad = gr.clone()
        a.step()
        b.step()
        assert torch.allclose(p, q, atol=2e-6), (it, float((p - q).abs().max()))


def test_dual_softmax_loss_kernel_vs_reference_autograd():
    """laff_dsl_forward_backward (loss.py:291-310) against the reference's value and autograd gradients."""
    from laff_b200 import loss as L
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dsl.npz"))
    for tag in ("small", "b128"):
        txt, vis = torch.from_numpy(d[tag + "/txt"]).cuda(), torch.from_numpy(d[tag + "/vis"]).cuda()
        loss, d_txt, d_vis = ops.dsl_forward_backward(txt, vis, 1000.0)
        ref = float(d[tag + "/loss"])
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (tag, loss.item(), ref)
        for got, key in ((d_txt, "/d_txt"), (d_vis, "/d_vis")):
            r = d[tag + key]
            assert np.abs(got.cpu().numpy() - r).max() <= 1e-4 * np.abs(r).max(), (tag, key)
        for temp in (1.0, 0.05):
            l2, a, b = ops.dsl_forward_backward(txt[:, 0].contiguous(), vis[:, 0].contiguous(), temp)
            r = float(d["%s/temp%g/loss" % (tag, temp)])
            assert abs(l2.item() - r) <= 2e-5 * abs(r)
            assert np.abs(a.cpu().numpy()[:, 0] - d["%s/temp%g/d_txt" % (tag, temp)]).max() <= 2e-4 * np.abs(d["%s/temp%g/d_txt" % (tag, temp)]).max()
        # module surface + autograd plumbing
        t = txt.clone().requires_grad_(True)
        out = L.DualSoftmaxLoss()(t, vis)
        out.backward()
        assert abs(out.item() - ref) <= 1e-5 * abs(ref) and torch.equal(t.grad, d_txt)


def test_gru_front_end_trains_with_the_model():
    """Token ids in, backward through time on the device: the LAFF model with the trainable GRU sentence encoder
    against three steps of the unmodified reference (its GruTxtEncoder under autograd), then graph replays."""
    from laff_b200 import text as T
    from test_train_cpu import check_params
    g, sd, H, steps = load_case("gru_rmsprop")
    D = int(g["meta"][1])
    dims = dict(SMALL, gru=int(g["gru_dim"]))
    c = cfg.laff_config(D, H, dims)
    c.dropout = 0.0
    c.optimizer, c.lr, c.grad_clip = str(g["optimizer"]), float(g["lr"]), float(g["grad_clip"])
    c.t2v_idx = T.IndexVec(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text", "vocab_gru.pkl"))
    c.we_dim, c.rnn_layer, c.we = int(g["we_dim"]), 1, None
    model = M.get_model("LAFF", torch.device("cuda"), c)
    load_numpy_state(model, sd)
    model.train()
    for s in range(steps + 4):
        k = s % steps
        vis_in = {str(n): g["step%d/vin/%s" % (k, n)] for n in g["vis_names"]}
        caps = {n: torch.from_numpy(g["step%d/tin/%s" % (k, n)]) for n in ("bow", "w2v", "clip")}
        caps["caption"] = [str(x) for x in g["step%d/captions" % k]]
        td = {"vis_feats": {n: torch.from_numpy(v) for n, v in vis_in.items()}, "captions": caps, "captions_task2": None,
              "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}
        loss = float(model(td, epoch=0)["triplet_loss"])
        if s >= steps:
            assert np.isfinite(loss)
            continue
        assert abs(loss - g["losses"][s]) <= (2e-5 if s == 0 else 5e-4) * abs(g["losses"][s]), (s, loss, g["losses"][s])
        if s == 0:
            grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
            ref_keys = [n[6:] for n in g.files if n.startswith("grad0/")]
            assert sorted(ref_keys) == sorted(grads.keys())
            for n in ref_keys:
                ref = g["grad0/" + n]
                got = grads[n].cpu().numpy().reshape(ref.shape)
                assert np.abs(got - ref).max() <= 1e-4 * max(1e-3, np.abs(ref).max()), (n, np.abs(got - ref).max())
        check_params({n: v.detach().cpu().numpy() for n, v in model.state_dict().items()}, g, s, float(g["lr"]), tight0=5e-5, frac=0.97)
    assert len(model._graphs) >= 1        # caption lengths differ between batches: one graph per shape
    model.eval()                          # and the eval path sees the trained GRU
    out = model.txt_net({"caption": ["a dog runs"], "bow": torch.zeros(1, SMALL["bow"]), "w2v": torch.zeros(1, SMALL["w2v"]),
                         "clip": torch.zeros(1, SMALL["clip"])})
    assert out.shape == (1, H, D // H) and bool(torch.isfinite(out).all())
