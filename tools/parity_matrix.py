"""Which 16-bit rounding moves ranks?  On the reference-trained fixture (tests/golden/trained_laff.npz): every combination
of projection-operand precision (x, W of the FC GEMMs) and embedding / similarity-operand type against the reference's
own ranks.  Prints one JSON line per combination; profiles/r02_parity_matrix.jsonl keeps the output."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from laff_b200 import ops, synth  # noqa: E402
from laff_b200.retrieval import GalleryIndex  # noqa: E402
from test_gpu_trained import build_model  # noqa: E402
from test_trained_fixture_cpu import load_trained  # noqa: E402


def main():
    g, sd, noise = load_trained()
    model, H = build_model(g, sd)
    for tag in ("c1", "c2"):
        n = int(g[tag + "/n"])
        vis, txt = synth.latent_collection(int(g[tag + "/seed"]), n, **noise)
        vin = {k: torch.from_numpy(x) for k, x in vis.items()}
        tin = {k: torch.from_numpy(x) for k, x in txt.items()}
        gt = torch.arange(n, device="cuda", dtype=torch.int32)
        ref_rank, ref_m = g[tag + "/rank0"], g[tag + "/metrics"]
        for proj in ("bf16x3", "fp16", "bf16"):
            for emb in ("split3", "fp16", "bf16"):
                dt = {"fp16": torch.float16, "bf16": torch.bfloat16, "split3": None}[emb]
                v32, v16 = model.vis_net.encode(vin, out16_dtype=dt, precision=proj)
                t32, t16 = model.txt_net.encode(tin, out16_dtype=dt, precision=proj)
                if emb == "split3":
                    q, gal = ops.split3_16(t32.reshape(n, -1), 0), ops.split3_16(v32.reshape(n, -1), 1)
                else:
                    q, gal = t16.reshape(n, -1), v16.reshape(n, -1)
                res = GalleryIndex(gal, n, H).search(q, gt, 10)
                r, m = res.rank0.cpu().numpy(), res.metrics.cpu().numpy()
                print(json.dumps({"case": tag, "projection_operands": proj, "similarity_operands": emb,
                                  "ranks_moved_pct": round(100 * float((r != ref_rank).mean()), 3),
                                  "max_abs_dRK": round(float(max(abs(m[i] - ref_m[i]) for i in range(3))), 3),
                                  "dMedR": float(m[3] - ref_m[3])}), flush=True)


if __name__ == "__main__":
    main()
