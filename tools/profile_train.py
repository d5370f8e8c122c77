"""Where one LAFF training step (B = 128) spends its time: CUDA-event time vs wall clock, and a per-kernel table.

    gpurun -- 'python tools/profile_train.py'
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laff_b200 import config as cfg, model as M, synth  # noqa: E402


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    c = cfg.laff_config(4096, 8, synth.DIMS)
    model = M.get_model("LAFF", torch.device(dev), c).train()
    B = 128
    td = {"vis_feats": {k: torch.randn(B, d, generator=g, device=dev) for k, d in c.vis_fc_layers[0].items()},
          "captions": {"gru": torch.randn(B, 1024, generator=g, device=dev), "bow": torch.zeros(B, 3981, device=dev),
                       "w2v": torch.randn(B, 500, generator=g, device=dev), "clip": torch.randn(B, 512, generator=g, device=dev)},
          "captions_task2": None, "vis_frame_feat_dict": {}, "vis_origin_frame_tuple": None}
    for _ in range(5):
        model(td)
    torch.cuda.synchronize()
    n = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        model(td)
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    print("per step: CUDA events %.3f ms, host issue %.3f ms, wall %.3f ms" % (e0.elapsed_time(e1) / n, t_issue / n * 1e3, t_wall / n * 1e3))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(10):
            model(td)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))


if __name__ == "__main__":
    main()
