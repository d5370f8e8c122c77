"""Times laff_fuse_forward (single-kernel fusion) in both MMA variants on the text and video LAFF nets.

    gpurun -- 'python tools/bench_fuse.py [rows]'   -> JSON lines, also appended to gpurun_out/fuse_bench.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laff_b200 import ops  # noqa: E402

NETS = {
    "txt (gru1024+bow3981+w2v500 FC, clip512 tiled)": ([1024, 3984, 504], 1),
    "vis (tf768+x3d2048+ircsn2048 FC, clip-ft512 tiled)": ([768, 2048, 2048], 1),
    "frame-vis (c3d2048+tf768+x3d2048+ircsn2048 FC+BN, frame512 tiled)": ([2048, 768, 2048, 2048], 1),
    # probes that separate the operand-feed limit from the epilogue limit (not reference configurations)
    "probe: one FC K=8192, no tiled feature (MMA-dominated)": ([8192], 0),
    "probe: four FC K=512, no tiled feature (epilogue-dominated)": ([512, 512, 512, 512], 0),
}


def timeit(fn, iters=7, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    quick = "--quick" in sys.argv  # one warm-up + one timed launch per variant: for runs under ncu (cycle counts)
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    out_path = os.path.join("gpurun_out", "fuse_bench.jsonl")
    os.makedirs("gpurun_out", exist_ok=True)
    results = {}
    for name, (ks, n_tiled) in NETS.items():
        fc = []
        for K in ks:
            fc.append({"x16": torch.randn(rows, K, generator=g, device=dev).to(torch.bfloat16),
                       "w16": (torch.randn(4096, K, generator=g, device=dev) * 0.02).to(torch.bfloat16),
                       "bias": torch.randn(4096, generator=g, device=dev) * 0.1, "activation": "tanh",
                       "bn_scale": torch.rand(4096, generator=g, device=dev) + 0.5,
                       "bn_shift": torch.randn(4096, generator=g, device=dev) * 0.1})
        tiled = [{"x": torch.randn(rows, 512, generator=g, device=dev),
                  "bn_scale": torch.rand(4096, generator=g, device=dev) + 0.5,
                  "bn_shift": torch.randn(4096, generator=g, device=dev) * 0.1} for _ in range(n_tiled)]
        aw = torch.randn(8, 512, generator=g, device=dev) / 22.6
        ab = torch.zeros(8, device=dev)
        flops = 2.0 * rows * 4096 * sum(ks)
        outs = {}
        for variant in (1, 2):
            ops.set_fuse_variant(variant)
            run = lambda: ops.fuse_forward(fc, tiled, aw, ab, 8, 512, want_f32=False, out16_dtype=torch.bfloat16)
            ms = timeit(run, iters=1, warmup=1) if quick else timeit(run)
            outs[variant] = run()[1].float()
            rec = {"kernel": "laff_fuse_forward cta_group::%d, %s, rows=%d" % (variant, name, rows), "ms": ms,
                   "achieved": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s"}
            print(json.dumps(rec), flush=True)
            with open(out_path, "a") as f:
                f.write(json.dumps(rec) + "\n")
        ops.set_fuse_variant(0)
        diff = (outs[1] - outs[2]).abs().max().item()
        print(json.dumps({"check": name, "max_abs_diff_variant1_vs_2": diff}), flush=True)
        results[name] = diff
        del fc, tiled, outs
        torch.cuda.empty_cache()
    assert all(v <= 1e-2 for v in results.values()), results


if __name__ == "__main__":
    main()
