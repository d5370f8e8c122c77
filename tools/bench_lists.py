"""Times GalleryIndex.ranked_lists (the writer lists of predictor.py:53-88) at gallery scale on one GPU.

    gpurun -- 'python tools/bench_lists.py [Q] [V] [k]'
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from laff_b200.retrieval import GalleryIndex  # noqa: E402


def main():
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 500
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(3)
    g16 = torch.empty(V, 4096, dtype=torch.bfloat16, device=dev)
    for s in range(0, V, 65536):
        x = torch.randn(min(65536, V - s), 8, 512, generator=gen, device=dev)
        g16[s:s + x.shape[0]] = (x / x.norm(dim=2, keepdim=True)).reshape(x.shape[0], -1).to(torch.bfloat16)
    x = torch.randn(Q, 8, 512, generator=gen, device=dev)
    q16 = (x / x.norm(dim=2, keepdim=True)).reshape(Q, -1).to(torch.bfloat16)
    idx = GalleryIndex(g16, V, 8)
    for chunk in (256, 1024, 2048):
        idx.ranked_lists(q16[:chunk], k, query_chunk=chunk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lv, li = idx.ranked_lists(q16, k, query_chunk=chunk)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"op": "ranked_lists", "Q": Q, "V": V, "k": k, "query_chunk": chunk, "ms": ms,
                          "queries_per_s": Q / (ms * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
