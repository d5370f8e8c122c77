"""Training step, CPU: the oracle's explicit forward / backward / optimizer restatement
against three real steps of the unmodified reference model (tests/golden/train_*.npz)."""
import os

import numpy as np
import pytest

from laff_b200 import synth
from oracle import laff_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def load_case(tag):
    g = np.load(os.path.join(HERE, "golden", "train_%s.npz" % tag))
    B, D, H, steps, seed, bn = [int(x) for x in g["meta"][:6]]
    sd = {k[4:]: g[k].copy() for k in g.files if k.startswith("sd0/")}
    return g, sd, H, steps


def step_inputs(g, s):
    names = [str(n) for n in g["vis_names"]]
    vis_in = {n: g["step%d/vin/%s" % (s, n)] for n in names}
    txt_in = {k: g["step%d/tin/%s" % (s, k)] for k in ("gru", "bow", "w2v", "clip")}
    return vis_in, txt_in


def check_params(sd, g, s, lr, tight0=3e-5, frac=0.99):
    """RMSprop / Adam normalise the gradient: an element whose gradient is at rounding-noise level (the shift-invariant
    logit bias, weights fed by ReLU zeros, ...) moves by up to lr / sqrt(1 - alpha) in a direction the noise decides, in
    the reference as much as here.  So: every element within that bound, and nearly all of them tight."""
    for k in sd:
        ref = g["sd%d/%s" % (s + 1, k)]
        err = np.abs(np.asarray(sd[k]).reshape(ref.shape).astype(np.float64) - ref)
        scale = max(1.0, np.abs(ref).max())
        assert err.max() <= 11 * lr * (s + 1) * scale, (k, s, err.max())
        if not k.endswith("embedding_common.0.bias"):
            tol = tight0 if s == 0 else 2e-4 * (s + 1)
            bad = int(np.sum(err > tol * scale))
            assert bad <= max(1, int((1 - frac) * err.size)), (k, s, bad, err.size)


def variant(g):
    """(with_ave, mul, loss kind) of a golden training case (older files: the shipped setting)."""
    m = [int(x) for x in g["meta"]]
    return (bool(m[7]), bool(m[8]), str(g["loss_kind"])) if len(m) > 8 else (False, False, "mrl")


@pytest.mark.parametrize("tag", ["rmsprop", "adam", "rmsprop_bn", "rmsprop_ave_mul", "adam_dsl"])
def test_oracle_train_steps_match_reference(tag):
    g, sd, H, steps = load_case(tag)
    state = {}
    opt, lr, clip = str(g["optimizer"]), float(g["lr"]), float(g["grad_clip"])
    with_ave, mul, loss_kind = variant(g)
    for s in range(steps):
        vis_in, txt_in = step_inputs(g, s)
        loss, grads, total = O.laff_train_step(sd, vis_in, txt_in, state, H, [synth.VIS_CLIP_FT], opt, lr, clip, with_ave=with_ave,
                                               mul=mul, loss_kind=loss_kind)
        assert abs(loss - g["losses"][s]) <= 2e-5 * abs(g["losses"][s]), (tag, s, loss, g["losses"][s])
        if s == 0:
            ref_keys = [k[6:] for k in g.files if k.startswith("grad0/")]
            assert sorted(ref_keys) == sorted(grads.keys())                     # same set of parameters receives gradients
            for k in ref_keys:
                ref = g["grad0/" + k]
                np.testing.assert_allclose(grads[k].reshape(ref.shape), ref, rtol=0, atol=2e-5 * max(1e-3, np.abs(ref).max()), err_msg=k)
        check_params(sd, g, s, lr)


def test_oracle_laff_ml_train_steps_match_reference():
    """LAFF-ml ('FrameLAFF'): frame-level attention in front of the video net, BatchNorm on every projected feature."""
    g, sd, H, steps = load_case("frame_rmsprop")
    ff = str(g["frame_feat"])
    state = {}
    lr, clip = float(g["lr"]), float(g["grad_clip"])
    for s in range(steps):
        vis_in, txt_in = step_inputs(g, s)
        loss, grads, total = O.laff_ml_train_step(sd, vis_in, g["step%d/frames" % s], ff, txt_in, state, H, str(g["optimizer"]), lr, clip)
        assert abs(loss - g["losses"][s]) <= 2e-5 * abs(g["losses"][s]), (s, loss, g["losses"][s])
        if s == 0:
            ref_keys = [k[6:] for k in g.files if k.startswith("grad0/")]
            assert sorted(ref_keys) == sorted(grads.keys())
            for k in ref_keys:
                ref = g["grad0/" + k]
                np.testing.assert_allclose(grads[k].reshape(ref.shape), ref, rtol=0, atol=2e-5 * max(1e-3, np.abs(ref).max()), err_msg=k)
        check_params(sd, g, s, lr)


@pytest.mark.parametrize("tag", ["frame_amp_rmsprop", "frame_amp_adam"])
def test_oracle_float16_branch_follows_reference_scaler(tag):
    """The reference's float16 branch (model/model.py:970-989: autocast + GradScaler, clip_grad_norm_ on the scaled
    gradients) over 12 steps of the unmodified reference (torch's CPU autocast / GradScaler bound to the names the
    reference imports, tests/golden/make_golden_train.py).  The restatement works on exact fp32 gradients and emulates
    the fp16 overflow on the parameter gradients: the loss-scale trajectory and the skipped steps must be IDENTICAL;
    losses agree to fp16 forward rounding; the clipped gradient of the first executed step has norm grad_clip / S."""
    g, sd, H, steps = load_case(tag)
    ff = str(g["frame_feat"])
    state, scaler = {}, {"scale": 65536.0, "tracker": 0}
    lr, clip, opt = float(g["lr"]), float(g["grad_clip"]), str(g["optimizer"])
    first = int(g["first_executed_step"])
    for s in range(steps):
        vis_in, txt_in = step_inputs(g, s)
        S = scaler["scale"]
        loss, grads, total = O.laff_ml_train_step(sd, vis_in, g["step%d/frames" % s], ff, txt_in, state, H, opt, lr, clip, scaler=scaler)
        assert scaler["scale"] == float(g["scales"][s]) and scaler["skipped"] == bool(g["skipped"][s]), (s, scaler, g["scales"][s])
        assert abs(loss - g["losses"][s]) <= 1e-2 * abs(g["losses"][s]), (s, loss, g["losses"][s])
        if s == first:
            n = np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for v in grads.values()))
            nref = np.sqrt(sum(float((g[k].astype(np.float64) ** 2).sum()) for k in g.files if k.startswith("grad_first/")))
            assert abs(n - clip / S) <= 1e-3 * clip / S and abs(nref - clip / S) <= 2e-2 * clip / S, (n, nref, clip / S)
            for k in grads:   # direction of the clipped gradient: fp16 autocast noise on the reference side
                ref = g["grad_first/" + k].astype(np.float64).ravel()
                got = grads[k].astype(np.float64).ravel()
                if np.linalg.norm(ref) > 1e-3 * nref:
                    cos = float(ref @ got / (np.linalg.norm(ref) * np.linalg.norm(got)))
                    assert cos >= 0.99, (k, cos)
    # parameters after 12 steps (9 executed): each executed step moves an element by at most ~lr / sqrt(1 - alpha)
    executed = int(steps - g["skipped"].sum())
    for k in sd:
        ref = g["sd%d/%s" % (steps, k)]
        err = np.abs(np.asarray(sd[k]).reshape(ref.shape).astype(np.float64) - ref)
        if "running_" in k or "num_batches" in k:
            continue
        assert err.max() <= 11 * lr * executed, (k, err.max())
        moved = np.abs(ref - g["sd0/" + k].reshape(ref.shape))
        # the big tensors (FC weights) must follow the reference's trajectory; the 1 x d_h logit weights and biases see
        # gradients at the level of the reference's fp16 rounding noise and only get the hard bound above
        if moved.max() > 0 and err.size >= 1024:
            assert np.median(err) <= 0.25 * max(np.median(moved), 1e-7), (k, np.median(err), np.median(moved))


def gru_tokens_of(g, s):
    """Token ids of step s's captions under the test vocabulary (IndexVec: <start> words <end>, unknown -> <unk>)."""
    from laff_b200 import text as T
    idx = T.IndexVec(os.path.join(HERE, "golden", "text", "vocab_gru.pkl"))
    return [idx.encoding(str(c)) for c in g["step%d/captions" % s]]


def test_oracle_gru_front_end_training_matches_reference():
    """The reference's real GruTxtEncoder (embedding + GRU, trained through autograd BPTT) inside the LAFF model."""
    g, sd, H, steps = load_case("gru_rmsprop")
    assert "txt_net.encoder.rnn_encoder.rnn.weight_hh_l0" in sd and "grad0/txt_net.encoder.rnn_encoder.we.weight" in g.files
    state = {}
    lr, clip = float(g["lr"]), float(g["grad_clip"])
    for s in range(steps):
        vis_in = {str(n): g["step%d/vin/%s" % (s, n)] for n in g["vis_names"]}
        txt_in = {k: g["step%d/tin/%s" % (s, k)] for k in ("bow", "w2v", "clip")}
        loss, grads, total = O.laff_train_step(sd, vis_in, txt_in, state, H, [synth.VIS_CLIP_FT], str(g["optimizer"]), lr, clip,
                                               gru_tokens=gru_tokens_of(g, s))
        assert abs(loss - g["losses"][s]) <= 2e-5 * abs(g["losses"][s]), (s, loss, g["losses"][s])
        if s == 0:
            ref_keys = [k[6:] for k in g.files if k.startswith("grad0/")]
            assert sorted(ref_keys) == sorted(grads.keys())
            for k in ref_keys:
                ref = g["grad0/" + k]
                np.testing.assert_allclose(grads[k].reshape(ref.shape), ref, rtol=0, atol=2e-5 * max(1e-3, np.abs(ref).max()), err_msg=k)
        check_params(sd, g, s, lr)
