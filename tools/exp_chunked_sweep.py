"""Experiment: one sweep of 10 000 queries vs the same queries in 2 / 4 / 8 launches (device-resident inputs, same box,
alternating order).  gpurun -- 'python tools/exp_chunked_sweep.py'"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from laff_b200 import synth  # noqa: E402
from laff_b200.retrieval import GalleryIndex  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    Q, V, H = 10000, 1000000, 8
    gen = torch.Generator(device=dev).manual_seed(1)
    q16 = bench.unit_rows(Q, gen, dev, torch.bfloat16)
    gt = ((torch.arange(Q, device=dev) * 97) % V).to(torch.int32)
    g16 = bench.build_gallery_shard(0, V, q16, gt.long(), synth.sigma_for_recall(V, 4096), dev)
    idx = GalleryIndex(g16, V, H)

    def run(per):
        outs = [idx.search(q16[s:s + per], gt[s:s + per], 10) for s in range(0, Q, per)]
        return torch.cat([o.rank0 for o in outs])

    ref = run(Q)
    for _ in range(2):
        run(Q)
    for rnd in range(3):
        for parts in (Q, 2500, 2560, 1280, 3334, 5000):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                r = run(parts)
            e1.record()
            torch.cuda.synchronize()
            assert torch.equal(r, ref)
            print("round %d: %d queries per launch: %.2f ms/step" % (rnd, parts, e0.elapsed_time(e1) / 8), flush=True)


if __name__ == "__main__":
    main()
