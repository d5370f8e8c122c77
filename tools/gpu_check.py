"""Bring-up battery for the CUDA kernels (run on the GPU box: ``python tools/gpu_check.py``).

Every section runs in its own subprocess under a timeout so that a trapped kernel cannot take the others down.
This is a developer tool; the parity tests proper live in tests/ and compare against oracle/.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _t():
    import torch
    return torch


def rand_emb(n, heads=8, dh=512, dtype=None, seed=0, device="cuda"):
    torch = _t()
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, heads, dh, generator=g, device=device)
    x = x / x.norm(dim=2, keepdim=True)
    return x.reshape(n, heads * dh).to(dtype or torch.bfloat16)


def sec_dense(cg, Q, V, D, dt):
    torch = _t()
    from laff_b200 import ops
    ops.set_tuning(cta_group=cg)
    dtype = torch.bfloat16 if dt == "bf16" else torch.float16
    g = torch.Generator(device="cuda").manual_seed(1)
    q = (torch.randn(Q, D, generator=g, device="cuda") / D ** 0.5).to(dtype)
    v = (torch.randn(V, D, generator=g, device="cuda") / D ** 0.5).to(dtype)
    out = ops.sim_dense(q, v, scale=0.125)
    torch.cuda.synchronize()
    ref = (q.double() @ v.double().T) * 0.125
    err = (out.double() - ref).abs()
    print("dense cg=%d Q=%d V=%d D=%d %s max_abs_err=%.3e ref_absmax=%.3e" % (cg, Q, V, D, dt, err.max().item(), ref.abs().max().item()))
    if err.max().item() > 1e-4:
        bad = (err > 1e-4)
        print("  bad fraction %.4f; bad rows %s; bad cols %s" % (
            bad.float().mean().item(), bad.any(1).nonzero().flatten()[:16].tolist(), bad.any(0).nonzero().flatten()[:16].tolist()))
        print("  out[0,:8]", out[0, :8].tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        raise SystemExit(1)


def sec_rank(cg, Q, V, D, k, chunk, mgroup):
    torch = _t()
    from laff_b200 import ops
    ops.set_tuning(cta_group=cg, chunk_tiles=chunk, m_group=mgroup)
    H = 8
    gal = rand_emb(V, H, D // H, seed=3)
    gt = (torch.arange(Q, device="cuda") * 97) % V
    noise = rand_emb(Q, H, D // H, dtype=torch.float32, seed=4)
    qf = gal[gt].float() * 0.6 + noise
    qf = qf.view(Q, H, -1)
    qf = (qf / qf.norm(dim=2, keepdim=True)).reshape(Q, D)
    q = qf.to(torch.bfloat16)
    # make some exact ties: duplicate gallery rows
    if V > 10:
        gal[5] = gal[int(gt[0])]
        gal[V - 1] = gal[int(gt[1])]
    dense = ops.sim_dense(q, gal, scale=0.125)
    r_ref, tv_ref, ti_ref = ops.rank_from_scores(dense, gt, k)
    sgt = ops.sim_gt_scores(q, gal, gt)
    torch.cuda.synchronize()
    sgt_dense = dense[torch.arange(Q, device="cuda"), gt]
    print("rank cg=%d Q=%d V=%d: sgt bit-equal to dense: %s (max diff %.3e)" % (
        cg, Q, V, bool((sgt * 0.125 == sgt_dense).all()), (sgt * 0.125 - sgt_dense).abs().max().item()))
    cnt, tv, ti = ops.sim_rank_topk(q, gal, sgt, gt, k, scale=0.125)
    torch.cuda.synchronize()
    ok_r = bool((cnt == r_ref).all())
    ok_i = bool((ti == ti_ref).all())
    ok_v = bool((tv == tv_ref).all())
    print("  rank equal %s, topk idx equal %s, topk val equal %s; R@1=%.2f" % (ok_r, ok_i, ok_v, (cnt == 0).float().mean().item() * 100))
    # cross-check rank_from_scores itself against torch on the dense matrix
    sg = dense[torch.arange(Q, device="cuda"), gt][:, None]
    cols = torch.arange(V, device="cuda")[None, :]
    beats = (dense > sg) | ((dense == sg) & (cols > gt[:, None]))
    beats &= cols != gt[:, None]
    r_t = beats.sum(1).to(torch.int32)
    print("  rank_from_scores vs torch: %s" % bool((r_t == r_ref).all()))
    m = ops.rank_metrics(cnt)
    torch.cuda.synchronize()
    import numpy as np
    rk = cnt.cpu().numpy().astype(np.float64)
    exp = [100.0 * (rk < 1).mean(), 100.0 * (rk < 5).mean(), 100.0 * (rk < 10).mean(), np.floor(np.median(rk)) + 1,
           rk.mean() + 1, (1.0 / (rk + 1)).mean()]
    print("  metrics dev %s" % [round(x, 6) for x in m.tolist()[:6]])
    print("  metrics np  %s" % [round(float(x), 6) for x in exp])
    if not (ok_r and ok_i and ok_v and bool((r_t == r_ref).all())):
        bad = (cnt != r_ref).nonzero().flatten()[:8].tolist()
        print("  first bad rank rows", bad, cnt[bad].tolist(), r_ref[bad].tolist())
        badk = (ti != ti_ref).any(1).nonzero().flatten()[:4].tolist()
        for b in badk:
            print("  row", b, ti[b].tolist(), ti_ref[b].tolist())
        raise SystemExit(1)


def sec_perf(cg, Q, V, chunk, mgroup, iters):
    torch = _t()
    from laff_b200 import ops
    ops.set_tuning(cta_group=cg, chunk_tiles=chunk, m_group=mgroup)
    D = 4096
    gal = torch.empty(V, D, dtype=torch.bfloat16, device="cuda")
    step = 65536
    for s in range(0, V, step):
        n = min(step, V - s)
        gal[s:s + n] = rand_emb(n, seed=100 + s)
    gt = (torch.arange(Q, device="cuda") * 97) % V
    q = rand_emb(Q, seed=7)
    sgt = ops.sim_gt_scores(q, gal, gt)
    ws = None
    for _ in range(2):
        out = ops.sim_rank_topk(q, gal, sgt, gt, 10, scale=0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = ops.sim_rank_topk(q, gal, sgt, gt, 10, scale=0.125)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * Q * V * D / (ms * 1e-3) / 1e12
    print("perf cg=%d chunk=%d mgroup=%d Q=%d V=%d: %.3f ms  %.1f TFLOP/s  %.0f q/s" % (cg, chunk, mgroup, Q, V, ms, tf, Q / (ms * 1e-3)))


def sec_perfx(kind, cg, Q, V, chunk, mgroup, ha, hb, iters=3, D=4096):
    """kind: null (mainloop + TMEM drain only) | rank | dense."""
    torch = _t()
    import ctypes as C
    from laff_b200 import ops, _capi
    ops.set_tuning(cta_group=cg, chunk_tiles=chunk, m_group=mgroup)
    gal = torch.empty(V, D, dtype=torch.bfloat16, device="cuda")
    step = 65536
    for s in range(0, V, step):
        n = min(step, V - s)
        gal[s:s + n] = rand_emb(n, 8, D // 8, seed=100 + s)
    gt = (torch.arange(Q, device="cuda") * 97) % V
    q = rand_emb(Q, 8, D // 8, seed=7)
    sink = torch.zeros(1024, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _capi.call("laff_debug_gemm", None, None, 0, 0, 0, 0, 0, 1, -1, ha, hb, None, st)  # set hints
    sgt = ops.sim_gt_scores(q, gal, gt)
    out = torch.empty(Q, V, device="cuda") if kind == "dense" else None

    def run():
        if kind == "null":
            _capi.call("laff_debug_gemm", C.c_void_p(q.data_ptr()), C.c_void_p(gal.data_ptr()), Q, V, D, D, D, 1, 1, ha, hb,
                       C.c_void_p(sink.data_ptr()), st)
        elif kind == "rank":
            ops.sim_rank_topk(q, gal, sgt, gt, 10, scale=0.125)
        else:
            ops.sim_dense(q, gal, 0.125, out=out)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * Q * V * D / (ms * 1e-3) / 1e12
    print("perfx %s cg=%d chunk=%d mgroup=%d hints=%d,%d Q=%d V=%d D=%d: %.3f ms  %.1f TFLOP/s  %.0f q/s" % (
        kind, cg, chunk, mgroup, ha, hb, Q, V, D, ms, tf, Q / (ms * 1e-3)))


def sec_fuse():
    torch = _t()
    from laff_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    B, H, dh = 300, 8, 512
    D = H * dh
    # l2norm_quantize
    x = torch.randn(B, D, device=dev)
    o = ops.l2norm_quantize(x, H, torch.float32)
    xr = x.view(B, H, dh)
    ref = (xr / (xr.pow(2).sum(2, keepdim=True).sqrt() + 1e-13 + 1e-14)).view(B, D)
    print("l2norm f32 err %.3e" % (o - ref).abs().max().item())
    ob = ops.l2norm_quantize(x, H, torch.bfloat16)
    print("l2norm bf16 equal to rounding of ref: %.4f" % (ob == ref.to(torch.bfloat16)).float().mean().item())
    # cast / split3
    xs = torch.randn(37, 500, device=dev)
    c = ops.cast_pad_16(xs, torch.bfloat16)
    print("cast_pad shape", tuple(c.shape), "equal", bool((c[:, :500] == xs.to(torch.bfloat16)).all()), "pad zero", bool((c[:, 500:] == 0).all()))
    l, r = ops.split3_16(xs, 0), ops.split3_16(xs, 1)
    approx = (l.double() @ r.double().T)
    exact = xs.double() @ xs.double().T
    print("split3 product rel err %.3e (plain bf16 %.3e)" % (
        ((approx - exact).abs().max() / exact.abs().max()).item(),
        ((c.double() @ c.double().T - exact).abs().max() / exact.abs().max()).item()))
    # project
    for K in (512, 504, 2048, 3984):
        xk = (torch.randn(B, K, device=dev)).to(torch.bfloat16)
        w = (torch.randn(D, K, device=dev) * (6.0 / (K + D)) ** 0.5).to(torch.bfloat16)
        b = torch.randn(D, device=dev) * 0.1
        sc = torch.rand(D, device=dev) + 0.5
        sh = torch.randn(D, device=dev) * 0.1
        for cg in (1, 2):
            ops.set_tuning(cta_group=cg)
            y = ops.project(xk, w, b, "tanh", sc, sh)
            torch.cuda.synchronize()
            ref = torch.tanh(xk.double() @ w.double().T + b.double()) * sc.double() + sh.double()
            print("project K=%d cg=%d err %.3e" % (K, cg, (y.double() - ref).abs().max().item()))
    # attention pool
    L = 4
    ys = [torch.tanh(torch.randn(B, D, device=dev)) for _ in range(L - 1)]
    xc = torch.randn(B, 512, device=dev)
    sc = torch.rand(D, device=dev) + 0.5
    sh = torch.randn(D, device=dev) * 0.1
    aw = torch.randn(H, dh, device=dev) / dh ** 0.5
    ab = torch.randn(H, device=dev) * 0.1
    for with_ave, mul in ((False, False), (True, False), (False, True), (True, True)):
        srcs = [{"x": xc, "bn_scale": sc, "bn_shift": sh}] + [{"y": y} for y in ys]
        out, out16, att = ops.attention_pool(srcs, aw, ab, H, dh, with_ave=with_ave, mul=mul, omega=0.6,
                                             out16_dtype=torch.bfloat16, want_att=True)
        torch.cuda.synchronize()
        yc = xc.repeat(1, H) * sc + sh
        Y = torch.stack([yc] + ys, 1).view(B, L, H, dh).double()
        mean = Y.mean(1, keepdim=True)
        common = Y * mean if mul else Y
        e = (common * aw.double()[None, None]).sum(3) + ab.double()[None, None]
        a = torch.softmax(e, 1)
        gsum = (a.unsqueeze(3) * Y).sum(1)
        if with_ave:
            gsum = gsum + 0.6 * mean.squeeze(1) * L
            a = a + 0.6 / L
        ref = gsum / (gsum.pow(2).sum(2, keepdim=True).sqrt() + 1e-14)
        print("pool with_ave=%d mul=%d err %.3e att err %.3e out16 ok %s" % (
            with_ave, mul, (out.double() - ref).abs().max().item(), (att.double() - a.permute(0, 2, 1)).abs().max().item(),
            bool((out16 == out.to(torch.bfloat16)).all())))
    # frame pool
    Bv, F, dim = 257, 32, 512
    fr = torch.randn(Bv, F, dim, device=dev)
    fr[:, 20:, :] = 0
    w = torch.randn(dim, device=dev) / dim ** 0.5
    for with_ave, mul in ((False, False), (True, True)):
        o = ops.frame_pool(fr, w, 0.05, with_ave=with_ave, mul=mul, omega=0.8)
        torch.cuda.synchronize()
        X = fr.double()
        mean = X.mean(1, keepdim=True)
        common = X * mean if mul else X
        e = (common * w.double()).sum(2) + 0.05
        a = torch.softmax(e, 1)
        gsum = (a.unsqueeze(2) * X).sum(1)
        if with_ave:
            gsum = gsum + 0.8 * mean.squeeze(1) * F
        ref = gsum / (gsum.pow(2).sum(1, keepdim=True).sqrt() + 1e-14)
        print("frame_pool with_ave=%d mul=%d err %.3e" % (with_ave, mul, (o.double() - ref).abs().max().item()))


def sec_fused():
    """Single-kernel fusion vs the two-kernel path vs fp64 torch, LAFF dims, several row counts."""
    torch = _t()
    import numpy as np
    from laff_b200 import config as cfg, loss as L, model as M, synth
    dev = "cuda"
    c = cfg.laff_config(4096, 8, synth.DIMS)
    for net_kind in ("txt", "vis"):
        net = (M.MultiScaleTxtEncoderAttention(c) if net_kind == "txt" else M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0]))
        sd = {k: torch.from_numpy(np.asarray(synth.param(5, k, tuple(v.shape)))).to(v.dtype).reshape(v.shape) for k, v in net.state_dict().items()}
        net.load_state_dict(sd)
        net = net.to(dev).eval()
        for rows in (1, 130, 1000, 2990):
            g = torch.Generator(device=dev).manual_seed(rows)
            if net_kind == "txt":
                feats = {"gru": torch.randn(rows, 1024, generator=g, device=dev), "bow": torch.randint(0, 3, (rows, 3981), generator=g, device=dev).float(),
                         "w2v": torch.randn(rows, 500, generator=g, device=dev), "clip": torch.randn(rows, 512, generator=g, device=dev)}
            else:
                feats = {k: torch.randn(rows, d, generator=g, device=dev).relu() for k, d in c.vis_fc_layers[0].items()}
            for prec in ("bf16", "bf16x3"):
                L.set_precision(prec)
                M.set_single_kernel_fusion(True)
                a32, a16 = net.encode(feats, out16_dtype=torch.bfloat16)
                M.set_single_kernel_fusion(False)
                b32, b16 = net.encode(feats, out16_dtype=torch.bfloat16)
                torch.cuda.synchronize()
                print("fused %s rows=%d %s: max|fused - two-kernel| = %.3e, out16 equal frac %.4f, norm err %.2e" % (
                    net_kind, rows, prec, (a32 - b32).abs().max().item(), (a16 == b16).float().mean().item(),
                    (a32.norm(dim=2) - 1).abs().max().item()))
    M.set_single_kernel_fusion(True)
    L.set_precision("bf16")
    # throughput: 65536 videos, LAFF vis net
    net = M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0]).to(dev).eval()
    rows = 65536
    g = torch.Generator(device=dev).manual_seed(1)
    feats = {k: torch.randn(rows, d, generator=g, device=dev).relu() for k, d in c.vis_fc_layers[0].items()}
    for single in (True, False):
        M.set_single_kernel_fusion(single)
        for _ in range(2):
            net.encode(feats, out16_dtype=torch.bfloat16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            net.encode(feats, out16_dtype=torch.bfloat16)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("encode %d videos single_kernel=%s: %.3f ms  %.1f TFLOP/s (39.85 MFLOP/video)  %.0f videos/s" % (
            rows, single, ms, rows * 39.85e6 / (ms * 1e-3) / 1e12, rows / (ms * 1e-3)))
    M.set_single_kernel_fusion(True)


def sec_fusedprof():
    torch = _t()
    import numpy as np
    from laff_b200 import config as cfg, model as M, synth
    dev = "cuda"
    c = cfg.laff_config(4096, 8, synth.DIMS)
    net = M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0]).to(dev).eval()
    rows = 32768
    g = torch.Generator(device=dev).manual_seed(1)
    feats = {k: torch.randn(rows, d, generator=g, device=dev).relu() for k, d in c.vis_fc_layers[0].items()}
    for _ in range(3):
        net.encode(feats, out16_dtype=torch.bfloat16)
    torch.cuda.synchronize()


def sec_loss():
    torch = _t()
    from laff_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    B, H, dh = 128, 8, 512
    vis = torch.randn(B, H, dh, device=dev)
    txt = vis + 0.8 * torch.randn(B, H, dh, device=dev)

    def ref_loss(txt, vis, margin, mv, direction, style):
        txt = txt.clone().double().requires_grad_(True)
        vis = vis.clone().double().requires_grad_(True)
        tot = 0
        for h in range(txt.shape[1]):
            s, im = txt[:, h], vis[:, h]
            sn = s / (s.pow(2).sum(1, keepdim=True).sqrt() + 1e-13 + 1e-14)
            imn = im / (im.pow(2).sum(1, keepdim=True).sqrt() + 1e-13 + 1e-14)
            sc = imn @ sn.t()
            d = sc.diag().view(-1, 1)
            I = torch.eye(B, device=dev) > .5
            cs = ci = None
            if direction in ("i2t", "bidir"):
                cs = (margin + sc - d.expand_as(sc)).clamp(min=0).masked_fill(I, 0)
            if direction in ("t2i", "bidir"):
                ci = (margin + sc - d.t().expand_as(sc)).clamp(min=0).masked_fill(I, 0)
            if mv:
                cs = cs.max(1)[0] if cs is not None else None
                ci = ci.max(0)[0] if ci is not None else None
            z = torch.zeros(1, device=dev, dtype=torch.double)
            cs = z if cs is None else cs
            ci = z if ci is None else ci
            tot = tot + (cs.sum() + ci.sum() if style == "sum" else cs.mean() + ci.mean())
        tot.backward()
        return tot.detach(), txt.grad, vis.grad

    for mv in (True, False):
        for direction in ("t2i", "i2t", "bidir"):
            for style in ("sum", "mean"):
                l, dt, dv = ops.mrl_forward_backward(txt, vis, 0.2, mv, direction, style)
                torch.cuda.synchronize()
                rl, rdt, rdv = ref_loss(txt, vis, 0.2, mv, direction, style)
                print("loss mv=%d %s %s: %.6f ref %.6f rel %.2e | dtxt rel %.2e dvis rel %.2e" % (
                    mv, direction, style, l.item(), rl.item(), abs(l.item() - rl.item()) / max(abs(rl.item()), 1e-12),
                    ((dt.double() - rdt).abs().max() / rdt.abs().max().clamp_min(1e-30)).item(),
                    ((dv.double() - rdv).abs().max() / rdv.abs().max().clamp_min(1e-30)).item()))
    sc = torch.randn(B, B, device=dev) * 0.1
    l, ds = ops.mrl_score_forward_backward(sc, 0.2, True, "t2i", "sum")
    s = sc.clone().double().requires_grad_(True)
    d = s.diag().view(-1, 1)
    I = torch.eye(B, device=dev) > .5
    ci = (0.2 + s - d.t().expand_as(s)).clamp(min=0).masked_fill(I, 0).max(0)[0].sum()
    ci.backward()
    print("score loss %.6f ref %.6f grad err %.2e" % (l.item(), ci.item(), (ds.double() - s.grad).abs().max().item()))


SECTIONS = {
    "dense_cg1_tiny": lambda: sec_dense(1, 128, 256, 64, "bf16"),
    "dense_cg1_k": lambda: sec_dense(1, 128, 256, 4096, "bf16"),
    "dense_cg1_odd": lambda: sec_dense(1, 300, 1000, 4096, "bf16"),
    "dense_cg1_f16": lambda: sec_dense(1, 300, 1000, 4096, "fp16"),
    "dense_cg1_big": lambda: sec_dense(1, 2990, 2990, 4096, "bf16"),
    "dense_cg2_tiny": lambda: sec_dense(2, 256, 256, 64, "bf16"),
    "dense_cg2_k": lambda: sec_dense(2, 256, 256, 4096, "bf16"),
    "dense_cg2_odd": lambda: sec_dense(2, 300, 1000, 4096, "bf16"),
    "dense_cg2_big": lambda: sec_dense(2, 2990, 2990, 4096, "bf16"),
    "dense_cg2_k520": lambda: sec_dense(2, 300, 1000, 520, "bf16"),
    "rank_cg1": lambda: sec_rank(1, 1000, 5000, 4096, 10, 4, 3),
    "rank_cg2": lambda: sec_rank(2, 1000, 5000, 4096, 10, 4, 3),
    "rank_cg2_b": lambda: sec_rank(2, 2990, 2990, 4096, 16, 16, 10),
    "rank_cg2_k1": lambda: sec_rank(2, 700, 9000, 4096, 1, 2, 2),
    "rank_cg2_c": lambda: sec_rank(2, 1000, 70001, 4096, 10, 16, 10),
    "fuse": sec_fuse,
    "fused": sec_fused,
    "fusedprof": sec_fusedprof,
    "loss": sec_loss,
    "perf_cg1": lambda: sec_perf(1, 10000, 200000, 16, 10, 3),
    "perf_cg2": lambda: sec_perf(2, 10000, 200000, 16, 10, 3),
    "perf_cg2_1m": lambda: sec_perf(2, 10000, 1000000, 16, 10, 2),
    "perf_cg1_1m": lambda: sec_perf(1, 10000, 1000000, 16, 10, 2),
}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--section":
        name = sys.argv[2]
        if name.startswith("perfx:"):
            a = name.split(":")
            sec_perfx(a[1], *[int(x) for x in a[2:]])
        else:
            SECTIONS[name]()
        return
    names = sys.argv[1:] or list(SECTIONS)
    failed = []
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--section", n], capture_output=True, text=True, timeout=300)
            out, rc = r.stdout + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = "TIMEOUT\n" + str(e.stdout)[-2000:], -9
        print("=== %s rc=%d (%.1fs)\n%s" % (n, rc, time.time() - t0, out), flush=True)
        if rc:
            failed.append(n)
    print("FAILED:", failed)


if __name__ == "__main__":
    main()
