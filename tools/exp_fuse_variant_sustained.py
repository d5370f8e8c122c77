"""Experiment: gallery fusion of 1 M videos (mode B's fusion leg) with the fused kernel forced to cta_group::1 / ::2,
several passes back to back so the GPU sits at its power cap.  gpurun -- 'python tools/exp_fuse_variant_sustained.py'"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from laff_b200 import ops  # noqa: E402
from laff_b200.retrieval import GalleryIndex  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    V = 1000000
    vis_net, dims = bench.build_vis_net(dev)
    raw = bench.raw_gallery_features(0, V, dims, dev)
    for rnd in range(3):
        for variant in (1, 2, 0):
            ops.set_fuse_variant(variant)
            GalleryIndex.from_features(vis_net, raw, V)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                idx = GalleryIndex.from_features(vis_net, raw, V)
            e1.record()
            torch.cuda.synchronize()
            del idx
            print("round %d variant %d (0 = auto): %.2f ms per 1 M videos" % (rnd, variant, e0.elapsed_time(e1) / 4), flush=True)
    ops.set_fuse_variant(0)


if __name__ == "__main__":
    main()
