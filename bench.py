#!/usr/bin/env python
"""bench.py — queries/sec ranked against a 1M-video gallery (BASELINE.json metric, config C5) on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU path (oracle port) on the host cores

A step = one pass of the hot path over one batch of synthetic queries: fuse the text features of Q = 10 000 queries
(gru + bow + w2v FC projections, CLIP tiled, LAFF pooling), then rank them against the V = 1 000 000-video gallery of
fused embeddings resident in HBM (tensor-core similarity sweep fused with exact rank counting and top-10, metrics on
device).  With N > 1 the gallery is sharded across the ranks (strong scaling: total work fixed) and the per-shard
results are merged with NCCL.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm (rank 0 only) is meant to use all host cores, and
# the BLAS behind numpy reads the variable when it is imported -- so this has to happen before the imports below.
if "reference" in sys.argv[1:] and os.environ.get("RANK", "0") == "0":
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)
# stdout carries exactly one JSON line.  NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION (set on some boxes,
# in the environment or an nccl.conf) and honours NCCL_DEBUG_FILE only above that level: raise VERSION to WARN and send
# the debug stream to stderr.
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402
import torch  # noqa: E402

Q_FULL, V_FULL, HEADS, HEAD_DIM, TOPK = 10000, 1000000, 8, 512, 10
D = HEADS * HEAD_DIM
METRIC = "queries/sec ranked vs 1M-video gallery"
UNIT = "queries/s"
FLOP_PER_PAIR = 2 * D            # 8192 FLOP per (query, video) similarity
FLOP_PER_QUERY_FUSE = 45.10e6    # 2*4096*(1024+3981+500)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor_tflops": float(d["bf16_tflops_sustained"]), "hbm_gbs": float(d["hbm_gbs"]), "source": "measured"}
    return {"tensor_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}  # B200_PROFILING.md fallback (sustained)


def sweep_launches(Q, V_local):
    """Kernel launches per laff_sim_rank_topk call: one per group of 10 row tiles (2560 queries) and per ~480 gallery
    column tiles (csrc/sim.cu: sweep_tiles_per_launch)."""
    groups = (Q + 2559) // 2560
    tiles = (V_local + 255) // 256
    split = max(1, tiles // 480)
    per = (tiles + split - 1) // split
    return groups * ((tiles + per - 1) // per)


def recorded_traffic():
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("rank_sweep_dram_bytes_per_launch_1gpu")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            busy = [s for s, w in zip(sm, power) if w >= 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


# --------------------------------------------------------------------------------------------------------------------
# synthetic workload
# --------------------------------------------------------------------------------------------------------------------
def build_txt_net(device):
    from laff_b200 import config as cfg, model as M, synth
    c = cfg.laff_config(D, HEADS, synth.DIMS)
    net = M.MultiScaleTxtEncoderAttention(c)
    sd = {k: torch.from_numpy(np.asarray(synth.param(1234, k, tuple(v.shape)))).to(v.dtype).reshape(v.shape)
          for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    return net.to(device).eval()


def build_vis_net(device):
    from laff_b200 import config as cfg, model as M, synth
    c = cfg.laff_config(D, HEADS, synth.DIMS)
    net = M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0])
    sd = {k: torch.from_numpy(np.asarray(synth.param(1234, k, tuple(v.shape)))).to(v.dtype).reshape(v.shape)
          for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    return net.to(device).eval(), dict(c.vis_fc_layers[0])


def raw_gallery_features(lo, hi, dims, device):
    """Raw fp32 video features of gallery rows [lo, hi) (mode B input, 21.5 KB/video): pooled-CNN-like (ReLU of a
    normal) for the backbone features, plain normal for the CLIP feature.  Resident in HBM like the reference's
    BigFile-backed features would be after loading."""
    from laff_b200 import synth
    n = hi - lo
    gen = torch.Generator(device=device).manual_seed(5150 + lo)
    out = {}
    for name, d in dims.items():
        x = torch.empty(n, d, dtype=torch.float32, device=device)
        for s in range(0, n, 131072):
            e = min(n, s + 131072)
            x[s:e].normal_(generator=gen)
        if name != synth.VIS_CLIP_FT:
            x.clamp_(min=0)
        out[name] = x
    return out


def query_features(Q, pinned: bool):
    """Synthetic per-encoder text features of Q queries on the host (optionally pinned).  The bag-of-words feature is
    what a tokenised caption is -- 8 vocabulary ids per query, CSR ('bow_csr': ops.SparseRows) -- not the dense
    [Q, 3981] count matrix the reference expands it to (txt2vec.py:56-63): 32 bytes instead of 15.9 KB per query."""
    g = torch.Generator().manual_seed(1234 + 5)
    from laff_b200 import ops, synth
    feats = {"gru": torch.randn(Q, synth.DIMS["gru"], generator=g),
             "w2v": torch.randn(Q, synth.DIMS["w2v"], generator=g),
             "clip": torch.randn(Q, synth.DIMS["clip"], generator=g)}
    ids = torch.randint(0, synth.DIMS["bow"], (Q, 8), generator=g)
    feats["bow_csr"] = ops.SparseRows(torch.arange(Q + 1, dtype=torch.int64) * 8, ids.reshape(-1).to(torch.int32), synth.DIMS["bow"])
    if pinned:
        feats = {k: v.pin_memory() for k, v in feats.items()}
    return feats


def feature_bytes(feats):
    return sum(v.nbytes() if hasattr(v, "nbytes") and callable(v.nbytes) else v.numel() * v.element_size() for v in feats.values())


def op_dtype(precision):
    return torch.float16 if precision == "fp16" else torch.bfloat16


def unit_rows(n, gen, device, dtype=torch.float32):
    x = torch.randn(n, HEADS, HEAD_DIM, generator=gen, device=device)
    return (x / x.norm(dim=2, keepdim=True)).reshape(n, D).to(dtype)


def build_gallery_shard(lo, hi, q_emb, gt, sigma, device, dtype=torch.float16):
    """Rows [lo, hi) of the synthetic gallery: unit-norm noise per head; row gt(i) = normalise(q_i + sigma * noise) so
    that R@1 is ~30 %.  Deterministic in the global row index, independent of the sharding."""
    n = hi - lo
    g16 = torch.empty(n, D, dtype=dtype, device=device)
    block = 65536
    for s in range(lo - lo % block, hi, block):
        gen = torch.Generator(device=device).manual_seed(9000 + s // block)
        rows = unit_rows(block, gen, device)
        a, b = max(s, lo), min(s + block, hi)
        g16[a - lo:b - lo] = rows[a - s:b - s].to(dtype)
    own = ((gt >= lo) & (gt < hi)).nonzero().flatten()
    if own.numel():
        gen = torch.Generator(device=device).manual_seed(777)
        noise = unit_rows(gt.numel(), gen, device)[own]
        planted = q_emb[own].float() + sigma * noise
        planted = planted.view(-1, HEADS, HEAD_DIM)
        planted = (planted / planted.norm(dim=2, keepdim=True)).reshape(-1, D)
        g16[gt[own] - lo] = planted.to(dtype)
    return g16


# --------------------------------------------------------------------------------------------------------------------
# arms
# --------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(q_emb, g_host, gt, sample, steps, warmup, threads):
    """The reference's evaluation path restated on the host (oracle port): per step `sample` queries against the whole
    gallery — get_txt2vis_matrix per 100k-video chunk, np.argsort per row, ground-truth search, metrics."""
    from oracle import laff_oracle as O
    torch.set_num_threads(threads)
    times = []
    Q = q_emb.shape[0]
    for it in range(warmup + steps):
        lo = (it * sample) % max(1, Q - sample + 1)
        t0 = time.perf_counter()
        O.retrieve_cpu(q_emb[lo:lo + sample], g_host, gt[lo:lo + sample], HEADS, k=TOPK, chunk=100000, threads=threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times


def staged_reference():
    """(model.model, evaluation) of the unmodified reference staged under baseline/_ref (oracle/stage_reference.py), or None."""
    try:
        from oracle import ref_loader
        return ref_loader.load() if ref_loader.available() else None
    except Exception as e:   # a broken staging must not take the arm down: the port below is the documented fallback
        print("bench.py: staged reference unusable (%s: %s); timing the oracle port" % (type(e).__name__, e), file=sys.stderr)
        return None


def reference_model(mm):
    """The reference's own LAFF model object (W2VVPP_MultiHeadAttention, configs.laff) at toy feature widths: the arm only
    calls its get_txt2vis_matrix (model/model.py:1003-1016), which has no parameters."""
    import importlib
    import types
    cfg = importlib.import_module("configs.laff").config()
    cfg.adjust_parm("0_12_0_12_0_0_1")
    cfg.vis_fc_layers = [{"clip_finetune_8frame_uniform_1103": HEAD_DIM, "X3D_L": 8}, D]
    cfg.txt_fc_layers = [0, D]
    cfg.multi_head_attention = {"dropout": 0.0, "heads": HEADS, "embed_dim_qkv": HEAD_DIM}
    cfg.clip_opt = dict(cfg.clip_opt, size=HEAD_DIM)
    cfg.rnn_size = 8
    cfg.t2v_bow = types.SimpleNamespace(ndims=8)
    cfg.t2v_w2v = types.SimpleNamespace(ndims=8)
    return mm.W2VVPP_MultiHeadAttention(cfg).eval()


def reference_cpu_steps(ref, q_emb, g_host, gt, sample, steps, warmup, threads):
    """The reference's OWN evaluation code on the host: per step `sample` queries against the whole gallery --
    W2VVPP.get_txt2vis_matrix (per-head loss.cosine_sim + cat + mean, model/model.py:1003-1016) over 100k-video gallery
    tiles as predict() tiles it (model/model.py:1060-1073), then get_predict_file's ranking loop verbatim
    (predictor.py:232-246: np.argsort of every row, string match of the ground-truth id in the re-ordered id array,
    0/1 label matrix) and evaluation.eval (evaluation.py:92-109)."""
    import contextlib
    mm, reval = ref
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(sys.stderr):    # the reference's config / constructors print; stdout carries one JSON line
        model = reference_model(mm)
    V = g_host.shape[0]
    vis_ids = ["video%d" % i for i in range(V)]
    times, last = [], None
    Q = q_emb.shape[0]
    for it in range(warmup + steps):
        lo = (it * sample) % max(1, Q - sample + 1)
        t0 = time.perf_counter()
        with torch.no_grad():
            txt = torch.from_numpy(q_emb[lo:lo + sample]).view(sample, HEADS, HEAD_DIM)
            t2i_matrix = np.empty((sample, V), dtype=np.float32)
            for c0 in range(0, V, 100000):
                vis = torch.from_numpy(g_host[c0:c0 + 100000]).view(-1, HEADS, HEAD_DIM)
                t2i_matrix[:, c0:c0 + vis.shape[0]] = model.get_txt2vis_matrix(txt, vis).numpy()
        txt_ids = ["video%d#enc#0" % int(g) for g in gt[lo:lo + sample]]
        inds = np.argsort(t2i_matrix, axis=1)                                        # predictor.py:232
        label_matrix = np.zeros(inds.shape)                                           # predictor.py:236
        for index in range(inds.shape[0]):                                            # predictor.py:239-244
            ind = inds[index][::-1]
            gt_index = np.where(np.array(vis_ids)[ind] == txt_ids[index].split('#')[0])[0]
            label_matrix[index][gt_index] = 1
        with contextlib.redirect_stdout(sys.stderr):
            last = reval.eval(label_matrix)                                           # predictor.py:246
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, last


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    V = args.videos
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    gen = torch.Generator(device=dev).manual_seed(4321)
    g_host = np.empty((V, D), dtype=np.float32)
    for s in range(0, V, 65536):
        n = min(65536, V - s)
        g_host[s:s + n] = unit_rows(n, gen, dev).cpu().numpy()
    Qr = max(sample * 2, 256)
    gt = (np.arange(Qr, dtype=np.int64) * 97) % V
    from laff_b200 import synth
    noise = unit_rows(Qr, gen, dev).cpu().numpy()
    q = synth.unit_heads(g_host[gt] + synth.sigma_for_recall(V, D) * noise, HEADS)
    ref = staged_reference()
    if ref is not None:
        times, _ = reference_cpu_steps(ref, q, g_host, gt, sample, args.steps, args.warmup, cores)
        kind = "reference"
        what = ("%d queries x %d videos per step through the UNMODIFIED reference staged under baseline/_ref: W2VVPP.get_txt2vis_matrix "
                "over 100k-video gallery tiles (model/model.py:1003-1016), get_predict_file's np.argsort + id-matching loop "
                "(predictor.py:232-246) and evaluation.eval (evaluation.py:92-109); torch + numpy on %d threads" % (sample, V, cores))
    else:
        times = cpu_reference_steps(q, g_host, gt, sample, args.steps, args.warmup, cores)
        kind = "port"
        what = ("%d queries x %d videos per step, gallery in 100k-video chunks, numpy/BLAS + threaded argsort (oracle port of "
                "model.py:1003-1016, predictor.py:232-244, evaluation.py:81-89)" % (sample, V))
    ms = 1e3 * sum(times) / len(times)
    val = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5: %d-query samples ranked against a %d-video gallery of fused LAFF embeddings "
                                   "(8 heads x 512), top-%d + rank + R@K/MedR" % (sample, V, TOPK),
                       "queries_per_step": sample, "gallery": V},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


@torch.no_grad()
def parity_block(txt_net, index, feats_dev, gt_dev, res, n, world, precision):
    """T2 parity of the benchmarked precision on the benchmark's own inputs: the first n queries are fused and ranked
    again with fp32-grade arithmetic -- 3-term bf16-split operands ('bf16x3') in the projections AND in the similarity
    sweep, the path whose ranks are identical to the fp32 reference on the reference-trained fixture
    (tests/test_gpu_trained.py) -- against the same 16-bit gallery (its values are exactly representable in the
    split).  Reports how many ground-truth ranks / top-10 lists the 16-bit step moved and the recall deltas."""
    import torch.distributed as dist
    from laff_b200 import ops
    n = min(n, gt_dev.numel())
    part = {k: v[:n] for k, v in feats_dev.items()}
    t32, _ = txt_net.encode(part, precision="bf16x3")
    q3 = ops.split3_16(t32.reshape(n, -1), 0)
    gt = gt_dev[:n].to(torch.int32)
    lo, hi, H = index.lo, index.hi, index.heads
    owned = (gt >= lo) & (gt < hi)
    rows = index.g16[torch.where(owned, gt - lo, torch.zeros_like(gt)).long()].float() if hi > lo else torch.zeros(n, D, device=gt.device)
    sgt = ops.sim_gt_scores(q3, ops.split3_16(rows, 1), torch.arange(n, device=gt.device, dtype=torch.int32))
    sgt = torch.where(owned, sgt, torch.zeros_like(sgt))
    if world > 1:
        dist.all_reduce(sgt)
    count = torch.zeros(n, dtype=torch.int32, device=gt.device)
    vals, idxs = [], []
    for s in range(0, hi - lo, 131072):
        e = min(hi - lo, s + 131072)
        c, tv, ti = ops.sim_rank_topk(q3, ops.split3_16(index.g16[s:e].float(), 1), sgt, gt, TOPK, scale=1.0 / H, col_offset=lo + s)
        count += c
        vals.append(tv)
        idxs.append(ti)
    if vals:
        tv, ti = ops.topk_merge(torch.stack(vals), torch.stack(idxs), TOPK)
    else:
        tv = torch.full((n, TOPK), float("-inf"), device=gt.device)
        ti = torch.full((n, TOPK), -1, dtype=torch.int32, device=gt.device)
    if world > 1:
        dist.all_reduce(count)
        av = [torch.empty_like(tv) for _ in range(world)]
        ai = [torch.empty_like(ti) for _ in range(world)]
        dist.all_gather(av, tv.contiguous())
        dist.all_gather(ai, ti.contiguous())
        tv, ti = ops.topk_merge(torch.stack(av), torch.stack(ai), TOPK)
    ref_rank, got_rank = count.cpu(), res.rank0[:n].cpu()
    m_ref, m_got = ops.rank_metrics(count).cpu().tolist(), ops.rank_metrics(res.rank0[:n].contiguous()).cpu().tolist()
    moved = int((ref_rank != got_rank).sum())
    near_top = (ref_rank < TOPK) | (got_rank < TOPK)          # the ranks R@1/5/10 depend on
    moved_top = int(((ref_rank != got_rank) & near_top).sum())
    deep = ref_rank[ref_rank != got_rank]
    lists = int((ti.cpu() != res.topk_idx[:n].cpu()).any(1).sum())
    return {"what": "first %d queries re-fused and re-ranked with fp32-grade arithmetic (bf16x3: 3-term split operands in the "
                    "projections and the sweep) against the same gallery; counts of queries the %s step answers differently"
                    % (n, precision),
            "queries": n, "ranks_moved": moved, "ranks_moved_frac": moved / max(1, n),
            "ranks_moved_within_top10": moved_top, "queries_within_top10": int(near_top.sum()),
            "median_reference_rank_of_moved": float(deep.float().median()) if deep.numel() else None,
            "note": "the synthetic gallery is chance-level noise around the planted ground truths: a ground truth that is not "
                    "retrieved sits hundreds of ranks deep among near-equal scores, where a 1e-5 score error moves it a few places; "
                    "the ranks that decide R@1/5/10 are the *_within_top10 figures (trained-checkpoint parity: tests/test_gpu_trained.py)",
            "max_rank_shift": int((ref_rank - got_rank).abs().max()) if n else 0, "top10_lists_differ": lists,
            "max_abs_score_diff": float((tv - res.topk_val[:n]).abs().max()),
            "d_r1": m_got[0] - m_ref[0], "d_r5": m_got[1] - m_ref[1], "d_r10": m_got[2] - m_ref[2], "d_medr": m_got[3] - m_ref[3],
            "recall_ref": {"r1": m_ref[0], "r5": m_ref[1], "r10": m_ref[2], "medr": m_ref[3]}}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from laff_b200 import _capi, ops, synth
    from laff_b200.retrieval import GalleryIndex, Retriever, shard_bounds
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    Q, V = args.queries, args.videos
    lib = _capi.lib()

    txt_net = build_txt_net(dev)
    feats_host = query_features(Q, pinned=True)
    feats_dev = {k: v.to(dev) for k, v in feats_host.items()}
    gt = ((torch.arange(Q, dtype=torch.int64) * 97) % V)
    gt_host = gt.to(torch.int32).pin_memory()
    gt_dev = gt.to(dev)
    dt16 = op_dtype(args.precision)
    with torch.no_grad():
        _, q16 = txt_net.encode(feats_dev, out16_dtype=dt16, precision=args.precision)
    sigma = synth.sigma_for_recall(V, D)
    # shard sizes: equal, or (--balance 1, W > 1) proportional to every GPU's sustained sweep rate measured here -- the
    # GPUs of one box settle at different clocks under the power cap and a step is as slow as its slowest shard
    weights = None
    if world > 1 and args.balance:
        from laff_b200.retrieval import calibrate_rank_weights
        weights = calibrate_rank_weights(dev, world, seconds=args.balance_seconds)
    lo, hi = shard_bounds(V, world, rank, weights)
    g16 = build_gallery_shard(lo, hi, q16.reshape(Q, D), gt_dev, sigma, dev, dt16)
    index = GalleryIndex(g16, V, HEADS, rank, world, weights=weights)
    retr = Retriever(txt_net, index)
    retr.reserve_sms, retr.side_max_ctas = args.reserve_sms, args.side_ctas
    pieces = args.pieces if args.pieces > 0 else None
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # A step = Retriever.submit: the batch goes through the three pipeline stages (fuse + gather | sweep | merge) on their
    # own streams, consecutive steps overlap stage-wise; every step's result is complete (device arm) / read back to the
    # host (e2e arm) inside the timed region.  --pipeline 0 times the serial Retriever.rank instead.
    def step_device():
        if args.pipeline:
            return retr.submit(feats_dev, gt_dev, TOPK, pieces=pieces, inputs_ready=False)
        return retr.rank(feats_dev, gt_dev, TOPK)

    def run_device(n):
        last = None
        for _ in range(n):
            last = step_device()
        return last.result() if args.pipeline else last      # steps complete in order: the last one's completion covers all

    def run_e2e(n):
        """n steps from pinned host buffers; each step's ranks / lists / metrics are read back with one packed D2H.  The
        pipelined loop reads step i's result after submitting step i + 1 (two steps in flight)."""
        out = prev = None
        for _ in range(n):
            if args.pipeline:
                cur = retr.submit(feats_host, gt_host, TOPK, pieces=pieces, fetch=True)
                if prev is not None:
                    out = prev.to_host()
                prev = cur
            else:
                out = retr.rank(feats_host, gt_host, TOPK, chunks=args.e2e_chunks).to_host()
        return prev.to_host() if args.pipeline else out

    # ---- device-resident inputs: `value` ------------------------------------------------------------------------
    res = run_device(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.laff_launch_count(1)
    index.timers = []
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    res = run_device(args.steps)
    t1.record()
    barrier()
    launches = int(lib.laff_launch_count(0))
    total_ms = t0.elapsed_time(t1)
    sweep_ms = [a.elapsed_time(b) for a, b in index.timers]
    index.timers = None
    sweeps_per_step = max(1, len(sweep_ms) // max(1, args.steps))
    # ---- host buffers through the public API: `e2e` -------------------------------------------------------------
    run_e2e(max(2, args.warmup // 2))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    out = run_e2e(args.steps)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the other 16-bit operand type on the same box, back to back (one GPU only): the price / gain of the parity choice
    other = None
    if world == 1 and args.ab_steps > 0:
        odt = torch.bfloat16 if dt16 == torch.float16 else torch.float16
        retr_o = Retriever(txt_net, GalleryIndex(g16.to(odt), V, HEADS, rank, world))
        retr_o.reserve_sms, retr_o.side_max_ctas = args.reserve_sms, args.side_ctas

        def run_o(n):
            last = None
            for _ in range(n):
                last = retr_o.submit(feats_dev, gt_dev, TOPK, pieces=pieces, inputs_ready=False) if args.pipeline else retr_o.rank(feats_dev, gt_dev, TOPK)
            return last.result() if args.pipeline else last
        run_o(3)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        ro = run_o(args.ab_steps)
        a1.record()
        barrier()
        o_ms = a0.elapsed_time(a1) / args.ab_steps
        om = ro.metrics.cpu().tolist()
        other = {"what": "the same step with %s operands (gallery converted from the %s one), %d steps right after the timed legs on the same "
                         "GPU: fp16 and bf16 share the nominal MMA rate, the power cap does not treat them alike; fp16 is the default "
                         "because it moves ~7x fewer ranks against the fp32 reference (tests/test_gpu_trained.py)"
                         % ("bf16" if odt == torch.bfloat16 else "fp16", args.precision, args.ab_steps),
                 "dtype": "bf16" if odt == torch.bfloat16 else "fp16", "ms_per_step": o_ms, "value": Q / (o_ms * 1e-3), "unit": UNIT,
                 "step_tflops": (Q * V * FLOP_PER_PAIR + Q * FLOP_PER_QUERY_FUSE) / (o_ms * 1e-3) / 1e12,
                 "step_frac_of_peak": (Q * V * FLOP_PER_PAIR + Q * FLOP_PER_QUERY_FUSE) / (o_ms * 1e-3) / 1e12 / peaks()["tensor_tflops"],
                 "recall": {"r1": om[0], "r5": om[1], "r10": om[2], "medr": om[3]}}
        del retr_o, ro
        torch.cuda.empty_cache()

    metrics = res.metrics.cpu().tolist()
    parity = parity_block(txt_net, index, feats_dev, gt_dev, res, args.parity_queries, world, args.precision) if args.parity_queries > 0 else None

    # ---- mode B: the gallery shard is re-fused from raw fp32 features inside the timed region -------
    modeb_ms = modeb_fuse_ms = 0.0
    if args.mode_b_steps > 0:
        del res
        vis_net, vis_dims = build_vis_net(dev)
        raw = raw_gallery_features(lo, hi, vis_dims, dev)

        def step_mode_b(timers=None):
            if timers is not None:
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record()
            idx_b = GalleryIndex.from_features(vis_net, raw, V, rank, world, out16_dtype=dt16)
            if timers is not None:
                b.record()
                timers.append((a, b))
            return Retriever(txt_net, idx_b).rank(feats_dev, gt_dev, TOPK)   # serial path: the shard is rebuilt every step

        step_mode_b()
        barrier()
        tb = []
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.mode_b_steps):
            step_mode_b(tb)
        t1.record()
        barrier()
        modeb_ms = t0.elapsed_time(t1) / args.mode_b_steps
        modeb_fuse_ms = sum(a.elapsed_time(b) for a, b in tb) / len(tb)
        del raw

    times = torch.tensor([total_ms, e2e_ms, sweeps_per_step * sum(sweep_ms) / max(1, len(sweep_ms)), modeb_ms, modeb_fuse_ms],
                         dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        mine = times[2:3].clone()
        allr = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = [round(float(x), 4) for x in allr.cpu()]
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, sweep_avg_ms, modeb_ms, modeb_fuse_ms = [float(x) for x in times.cpu()]

    if rank != 0:
        return
    pk = peaks()
    ms_per_step = total_ms / args.steps
    value = Q / (ms_per_step * 1e-3)
    e2e_value = Q / (e2e_ms / args.steps * 1e-3)
    n_local = hi - lo
    # per-GPU rate: the average shard (V / W videos) over the slowest rank's sweep time
    achieved = Q * (V / world) * FLOP_PER_PAIR / (sweep_avg_ms * 1e-3) / 1e12
    # summed over the ranks: every rank copies its 1/W slice of the query features and the whole ground-truth vector
    h2d = feature_bytes(feats_host) + world * gt_host.numel() * 4
    d2h = world * (Q * 4 + Q * TOPK * 8 + 8 * 8)   # every rank reads the (replicated) ranks, top-k lists and metrics back
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "C5: %d queries (gru1024+w2v500 FC -> 4096, bow3981 as CSR token ids -> gather-sum FC, CLIP512 tiled, LAFF pooling, 8x512) "
                               "ranked against a %d-video gallery of fused embeddings resident in HBM: similarity "
                               "sweep + exact rank + top-%d + R@K/MedR" % (Q, V, TOPK),
                   "queries": Q, "gallery": V, "gallery_per_gpu": n_local, "topk": TOPK,
                   "sharding": "gallery rows / %d" % world if weights is None else
                               "gallery rows over %d ranks in proportion to each GPU's sustained sweep rate (calibrated at start-up, "
                               "%.1f s per rank): fractions %s" % (world, args.balance_seconds, [round(w, 4) for w in weights]),
                   "l2": "inputs exceed L2 (gallery shard %.1f GB)" % (n_local * D * 2 / 1e9),
                   "recall": {"r1": metrics[0], "r5": metrics[1], "r10": metrics[2], "medr": metrics[3]},
                   "other_precision": other,
                   "pipeline": None if not args.pipeline else {
                       "what": "Retriever.submit: per piece of the batch fuse+all-gather+ground-truth scores | sweep | merge on three "
                               "streams (collectives of the side stages on their own communicators); stages of consecutive "
                               "pieces and steps overlap, every step's result completes inside the timed region",
                       "pieces_per_step": sweeps_per_step, "reserve_sms": args.reserve_sms, "side_ctas": args.side_ctas},
                   "mode_b": None if args.mode_b_steps <= 0 else {
                       "what": "same step with the gallery shard re-fused from raw fp32 video features (tf768+x3d2048+"
                               "ircsn2048 FC, clip-ft512 tiled; 21.5 KB/video resident in HBM) inside the timed region",
                       "value": Q / (modeb_ms * 1e-3), "unit": UNIT, "ms_per_step": modeb_ms, "steps": args.mode_b_steps,
                       "gallery_fusion_ms": modeb_fuse_ms,
                       "gallery_fusion_tflops": n_local * 2.0 * D * (768 + 2048 + 2048) / (modeb_fuse_ms * 1e-3) / 1e12}},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "parity": parity,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm_kernel<2, EpiRank<16>> (similarity sweep + rank + top-k; one sweep of the step's %d "
                                                  "queries = %d laff_sim_rank_topk call(s) = %d back-to-back launches of this kernel; CUDA "
                                                  "events on the sweep stream around every call, summed per step%s)"
                                                  % (Q, sweeps_per_step, sweep_launches(Q, n_local),
                                                     "; %d of the SMs are left to the pipeline's side streams while it runs" % args.reserve_sms
                                                     if args.pipeline and args.reserve_sms > 0 else ""),
                     "achieved": achieved, "peak": pk["tensor_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tensor_tflops"],
                     "peak_source": pk["source"] + " (bf16 dense sustained; fp16 and bf16 operands share the kind::f16 MMA rate)",
                     "flop_per_launch": Q * (V / world) * FLOP_PER_PAIR, "avg_launch_ms": sweep_avg_ms,
                     "avg_launch_ms_per_rank": per_rank,
                     "traffic": recorded_traffic() if world == 1 and V == V_FULL and Q == Q_FULL else None,
                     "traffic_source": "recorded: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this "
                                       "sweep (profiles/roofline_traffic.json), not measured in this run"},
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample
        g_host = np.empty((V, D), dtype=np.float32)
        for s in range(0, V, 131072):
            g_host[s:s + 131072] = g16[s:s + 131072].float().cpu().numpy()
        q_host = q16.reshape(Q, D)[: max(2 * sample, 256)].float().cpu().numpy()
        ref = staged_reference()
        if ref is not None:
            tms, ref_metrics = reference_cpu_steps(ref, q_host, g_host, gt.numpy()[: len(q_host)], sample, 2, 1, cores)
            kind, how = "reference", ("the unmodified reference staged under baseline/_ref: get_txt2vis_matrix + predictor.py:232-246 "
                                      "argsort / id-matching loop + evaluation.eval")
        else:
            tms = cpu_reference_steps(q_host, g_host, gt.numpy()[: len(q_host)], sample, 2, 1, cores)
            kind, how = "port", "oracle port of the reference's sim + argsort + metrics path"
        cval = sample / (sum(tms) / len(tms))
        line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": "%d queries x %d videos per step (2 steps after 1 warm-up), same embeddings as "
                                          "the GPU arm, %s" % (sample, V, how)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=Q_FULL)
    ap.add_argument("--videos", type=int, default=V_FULL)
    ap.add_argument("--cpu-sample", type=int, default=16,
                    help="queries per CPU-baseline step (the reference's own ranking loop costs ~0.7 s per query against 1 M "
                         "videos on 8 cores: 16 queries keep a step near 10 s and a 25-step reference run within minutes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--balance", type=int, default=0, help="1: size the gallery shards by each GPU's measured sweep rate (W > 1)")
    ap.add_argument("--balance-seconds", type=float, default=2.0)
    ap.add_argument("--pipeline", type=int, default=1, help="1: Retriever.submit (staged, overlapping steps); 0: serial Retriever.rank")
    ap.add_argument("--pieces", type=int, default=0,
                    help="query pieces per step of the pipelined path (0 = one per 2560 queries, at most 4 on one GPU and 2 with several)")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="SMs the sweep leaves to the pipeline's side streams (-1 = 0 on one GPU, where the side stages are "
                         "short kernels that fill the gaps between sweep launches, 4 with several ranks, where NCCL kernels wait "
                         "for their peers and must not hold SMs the sweep's persistent CTAs were sized for)")
    ap.add_argument("--side-ctas", type=int, default=2, help="CTAs per NCCL kernel on the side communicators")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit operand type of the projections and the similarity sweep (same MMA rate; fp16 moves ~7x fewer "
                         "ranks against the fp32 reference, tests/test_gpu_trained.py)")
    ap.add_argument("--ab-steps", type=int, default=5,
                    help="timed steps of the same workload with the other 16-bit operand type, reported as config.other_precision "
                         "(one GPU only; 0 = skip)")
    ap.add_argument("--parity-queries", type=int, default=256,
                    help="queries of the parity block: re-ranked with fp32-grade (bf16x3) fusion + sweep and compared (0 = skip)")
    ap.add_argument("--mode-b-steps", type=int, default=2, help="timed steps of the mode-B leg (0 = skip it)")
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="query pieces whose H2D copies overlap the sweep in the e2e leg (0 = by world size: 4 / 2 / 1 / 1 "
                         "at 1 / 2 / 4 / 8 GPUs -- a piece costs a fixed ~0.7 ms of launches and collectives, worth it "
                         "only while the per-rank copy is a visible share of the step)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.e2e_chunks <= 0:
        args.e2e_chunks = max(1, 4 // max(1, int(os.environ.get("WORLD_SIZE", "1"))))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.reserve_sms < 0:
        args.reserve_sms = 0 if world == 1 else 4
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
