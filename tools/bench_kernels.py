"""Per-kernel measurements for the §8 rows other than the headline sweep (run on the GPU box).

    python tools/bench_kernels.py            # prints one JSON line per kernel

Each entry: algorithmic bytes or FLOPs per launch / CUDA-event time -> achieved, against MEASURED_PEAKS.json.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from laff_b200 import config as cfg, loss as L, model as M, ops, synth  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), float(d["hbm_gbs"]), "measured"
    return 1400.0, 6650.0, "fallback"


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def emit(name, ms, flops=None, bytes_=None, note=""):
    tf, gb, src = peaks()
    rec = {"kernel": name, "ms": ms, "note": note, "peak_source": src}
    if flops is not None:
        a = flops / (ms * 1e-3) / 1e12
        rec.update(bound="tensor", achieved=a, peak=tf, unit="TFLOP/s", frac=a / tf)
    if bytes_ is not None:
        a = bytes_ / (ms * 1e-3) / 1e9
        rec.update(bound="hbm", achieved=a, peak=gb, unit="GB/s", frac=a / gb)
    print(json.dumps(rec), flush=True)


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > L2

    # F1 projection GEMM (+ bias, tanh, BN): 65536 rows, the four FC input widths
    rows = 65536
    for K in (2048, 768, 1024, 3984, 504):
        x16 = torch.randn(rows, K, generator=g, device=dev).to(torch.bfloat16)
        w16 = (torch.randn(4096, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
        b = torch.randn(4096, generator=g, device=dev) * 0.1
        sc = torch.rand(4096, generator=g, device=dev) + 0.5
        sh = torch.randn(4096, generator=g, device=dev) * 0.1
        y = torch.empty(rows, 4096, device=dev)
        ms = timeit(lambda: ops.project(x16, w16, b, "tanh", sc, sh, out=y), flush=flush)
        emit("laff_project K=%d rows=%d (bias+tanh+BN fused, fp32 out)" % (K, rows), ms, flops=2.0 * rows * K * 4096,
             note="writes %.2f GB of fp32 Y" % (rows * 4096 * 4 / 1e9))

    # F5/F6 pooling: L = 4 (3 projected + 1 tiled), fp32 in, fp32 + bf16 out
    ys = [torch.tanh(torch.randn(rows, 4096, generator=g, device=dev)) for _ in range(3)]
    xc = torch.randn(rows, 512, generator=g, device=dev)
    sc = torch.rand(4096, generator=g, device=dev) + 0.5
    sh = torch.randn(4096, generator=g, device=dev) * 0.1
    aw = torch.randn(8, 512, generator=g, device=dev) / 22.6
    ab = torch.zeros(8, device=dev)
    srcs = [{"x": xc, "bn_scale": sc, "bn_shift": sh}] + [{"y": t} for t in ys]
    ms = timeit(lambda: ops.attention_pool(srcs, aw, ab, 8, 512, out16_dtype=torch.bfloat16), flush=flush)
    emit("laff_attention_pool L=4 rows=%d" % rows, ms, bytes_=rows * (3 * 4096 * 4 + 512 * 4 + 4096 * 4 + 4096 * 2))

    # F7 frame pooling: 32 frames x 512 per video
    Bv = 65536
    fr = torch.randn(Bv, 32, 512, generator=g, device=dev)
    w = torch.randn(512, generator=g, device=dev) / 22.6
    ms = timeit(lambda: ops.frame_pool(fr, w, 0.0), flush=flush)
    emit("laff_frame_pool F=32 dim=512 videos=%d" % Bv, ms, bytes_=Bv * (32 * 512 * 4 + 512 * 4))

    # S1 normalise + quantise
    e = torch.randn(rows * 4, 4096, generator=g, device=dev)
    ms = timeit(lambda: ops.l2norm_quantize(e, 8, torch.bfloat16), flush=flush)
    emit("laff_l2norm_quantize rows=%d" % (rows * 4), ms, bytes_=rows * 4 * 4096 * 6)

    # operand preparation: fp32 features -> 16-bit TMA operands (mode B casts 19.5 KB per video), 3-term split (training)
    xc = torch.randn(131072, 2048, generator=g, device=dev)
    scratch = torch.empty(131072, 2048, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.cast_pad_16(xc, torch.bfloat16, out=scratch), flush=flush)
    emit("laff_cast_pad_16 rows=131072 K=2048", ms, bytes_=131072 * 2048 * 6)
    xw = torch.randn(4096, 3981, generator=g, device=dev)
    ms = timeit(lambda: ops.split3_16(xw, 1, torch.bfloat16), flush=flush)
    emit("laff_split3_16 rows=4096 K=3981 (weights of one training step)", ms, bytes_=4096 * (3981 * 4 + 3984 * 6))
    del xc, scratch, xw

    # L1/L2 loss step (C3)
    txt = torch.randn(128, 8, 512, generator=g, device=dev)
    vis = torch.randn(128, 8, 512, generator=g, device=dev)
    ms = timeit(lambda: ops.mrl_forward_backward(txt, vis, 0.2, True, "t2i", "sum"), iters=50)
    emit("laff_mrl_forward_backward B=128 H=8 d=512 (fwd+bwd)", ms, note="latency-bound; 134 MFLOP fwd")

    # C2: MV-test3k-shaped fused eval, 2990 x 2990, raw features in -> metrics out
    n = 2990
    c = cfg.laff_config(4096, 8, synth.DIMS)
    vis_net = M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0]).to(dev).eval()
    txt_net = M.MultiScaleTxtEncoderAttention(c).to(dev).eval()
    vin = {k: torch.randn(n, d, generator=g, device=dev) for k, d in c.vis_fc_layers[0].items()}
    tin = {"gru": torch.randn(n, 1024, generator=g, device=dev), "bow": torch.zeros(n, 3981, device=dev),
           "w2v": torch.randn(n, 500, generator=g, device=dev), "clip": torch.randn(n, 512, generator=g, device=dev)}
    gt = torch.arange(n, device=dev, dtype=torch.int32)
    from laff_b200.retrieval import GalleryIndex

    def c2():
        with torch.no_grad():
            _, v16 = vis_net.encode(vin, out16_dtype=torch.bfloat16)
            _, t16 = txt_net.encode(tin, out16_dtype=torch.bfloat16)
            return GalleryIndex(v16.reshape(n, -1), n, 8).search(t16.reshape(n, -1), gt, 10)
    ms = timeit(c2, iters=20)
    emit("C2 fused eval 2990x2990 (encode both sides + sweep + rank + top-10 + metrics)", ms,
         note="%.0f queries/s; latency-bound config" % (n / (ms * 1e-3)))

    def c2_dense():
        with torch.no_grad():
            _, v16 = vis_net.encode(vin, out16_dtype=torch.bfloat16)
            _, t16 = txt_net.encode(tin, out16_dtype=torch.bfloat16)
            s = ops.sim_dense(t16.reshape(n, -1), v16.reshape(n, -1), 0.125)
            r, tv, ti = ops.rank_from_scores(s, gt, 10)
            return ops.rank_metrics(r)
    ms = timeit(c2_dense, iters=20)
    emit("C2 via dense matrix 2990x2990 (encode + sim_dense + rank_from_scores + metrics)", ms,
         note="%.0f queries/s" % (n / (ms * 1e-3)))

    # C3 / T1: one full training step of the LAFF model, B = 128 (forward both nets in train mode, loss, backward,
    # clip + RMSprop), device-resident inputs; dropout 0.2 as shipped
    Bt = 128
    model = M.get_model("LAFF", torch.device(dev), c).train()
    lib = __import__("laff_b200._capi", fromlist=["lib"]).lib()
    td = {"vis_feats": {k: torch.randn(Bt, d, generator=g, device=dev) for k, d in c.vis_fc_layers[0].items()},
          "captions": {"gru": torch.randn(Bt, 1024, generator=g, device=dev), "bow": torch.zeros(Bt, 3981, device=dev),
                       "w2v": torch.randn(Bt, 500, generator=g, device=dev), "clip": torch.randn(Bt, 512, generator=g, device=dev)},
          "captions_task2": None, "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}
    model(td)
    lib.laff_launch_count(1)
    model(td)
    launches = int(lib.laff_launch_count(0))
    ms = timeit(lambda: model(td), iters=30)
    n_par = sum(p.numel() for p in model.parameters())
    emit("LAFF training step B=128 (train-mode forward, loss, backward, clip + RMSprop; %.1f M parameters)" % (n_par / 1e6), ms,
         note="%d kernel launches of this library per step; latency-bound (14.6 GFLOP of GEMM work)" % launches)

    # N2: text front-end at query-batch scale: 10 000 captions of 4..16 words over a 4 096-word vocabulary
    import tempfile
    from laff_b200 import text as T
    from laff_b200.bigfile import write_bigfile
    rs = np.random.RandomState(0)
    words = ["w%04d" % i for i in range(4096)]
    caps = [" ".join(rs.choice(words, size=rs.randint(4, 17))) for _ in range(10000)]
    vocab_b, vocab_g = T.Vocabulary("bow"), T.Vocabulary("gru")
    for w in ["<pad>", "<start>", "<end>", "<unk>"]:
        vocab_g.add(w)
    for w in words:
        vocab_b.add(w)
        vocab_g.add(w)
    with tempfile.TemporaryDirectory() as td:
        write_bigfile(td, words, rs.standard_normal((4096, 500)).astype(np.float32))
        bow, w2v = T.BowVec(None, vocab=vocab_b), T.W2Vec(td)
        idx = T.IndexVec(None, vocab=vocab_g)
        import types
        opt = types.SimpleNamespace(t2v_idx=idx, rnn_layer=1, we_dim=500, rnn_size=1024, pooling="mean", we=None, t2v_bow=bow, t2v_w2v=w2v)
        gru = M.GruTxtEncoder(opt).to(dev).eval()
        import time
        t0 = time.perf_counter()
        tok = [idx.encoding(c_) for c_ in caps[:1000]]
        py_ms = (time.perf_counter() - t0) * 1e4          # the reference's per-caption Python path, extrapolated to 10 000
        idx.encoding_batch(caps[:10])
        t0 = time.perf_counter()
        idx.encoding_batch(caps)
        host_ms = (time.perf_counter() - t0) * 1e3
        for name, fn in (("BoW counts (laff_bow_counts)", lambda: bow.encode_batch(caps)), ("word2vec means (laff_gather_mean)", lambda: w2v.encode_batch(caps)),
                         ("GRU 500->1024, mean pooling (embedding + %d-step recurrence on the GEMM engine)" % int(idx.encoding_batch(caps)[1].max()),
                          lambda: gru({"caption": caps}))):
            ms = timeit(fn, iters=5, warmup=2)
            emit("text front-end, 10000 captions: " + name, ms,
                 note="includes native tokenisation + lookup (%.1f ms for the GRU's token ids; %.0f ms through the per-caption Python path)"
                      % (host_ms, py_ms))


if __name__ == "__main__":
    main()
