"""Gallery fusion (mode B) alone: 1 M videos fused from raw fp32 features, with / without the cast overlap.
    python tools/bench_modeb.py [overlap_sms ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from laff_b200 import model as M  # noqa: E402
from laff_b200.retrieval import GalleryIndex  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    V = 1000000
    vis_net, dims = bench.build_vis_net(dev)
    raw = bench.raw_gallery_features(0, V, dims, dev)
    ref = None
    for sms in [int(a) for a in sys.argv[1:]] or [0, 12]:
        M.set_cast_overlap(sms)
        for _ in range(2):
            idx = GalleryIndex.from_features(vis_net, raw, V)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 4
        for _ in range(n):
            idx = GalleryIndex.from_features(vis_net, raw, V)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        same = True if ref is None else bool(torch.equal(ref, idx.g16))
        if ref is None:
            ref = idx.g16.clone()
        print("cast overlap on %2d SMs: %.2f ms per 1 M videos = %.0f TFLOP/s, output identical to serial: %s" % (
            sms, ms, V * 2.0 * 4096 * (768 + 2048 + 2048) / (ms * 1e-3) / 1e12, same), flush=True)


if __name__ == "__main__":
    main()
