"""Cycle breakdown of the fused kernel's epilogue phases (needs a library built with LAFF_NVCC_EXTRA=-DLAFF_FUSE_PROFILE).

    gpurun -- 'LAFF_NVCC_EXTRA=-DLAFF_FUSE_PROFILE python -m laff_b200.build && python tools/profile_fuse_phases.py'
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from laff_b200 import _capi, ops  # noqa: E402
from bench_fuse import NETS  # noqa: E402

NAMES = ["wait accumulator", "pass A", "post", "collect (exchange wait)", "pass B", "first tiled feature", "normalise + store", "features"]


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    lib = _capi.lib()
    buf = (C.c_ulonglong * 8)()
    for name, (ks, n_tiled) in NETS.items():
        fc = [{"x16": torch.randn(rows, K, generator=g, device=dev).to(torch.bfloat16),
               "w16": (torch.randn(4096, K, generator=g, device=dev) * 0.02).to(torch.bfloat16),
               "bias": torch.randn(4096, generator=g, device=dev) * 0.1, "activation": "tanh",
               "bn_scale": torch.rand(4096, generator=g, device=dev) + 0.5, "bn_shift": torch.randn(4096, generator=g, device=dev) * 0.1} for K in ks]
        tiled = [{"x": torch.randn(rows, 512, generator=g, device=dev), "bn_scale": torch.rand(4096, generator=g, device=dev) + 0.5,
                  "bn_shift": torch.randn(4096, generator=g, device=dev) * 0.1} for _ in range(n_tiled)]
        aw = torch.randn(8, 512, generator=g, device=dev) / 22.6
        ab = torch.zeros(8, device=dev)
        for variant in (1, 2):
            ops.set_fuse_variant(variant)
            ops.fuse_forward(fc, tiled, aw, ab, 8, 512, want_f32=False, out16_dtype=torch.bfloat16)
            torch.cuda.synchronize()
            lib.laff_debug_fuse_profile(buf, 1)
            ops.fuse_forward(fc, tiled, aw, ab, 8, 512, want_f32=False, out16_dtype=torch.bfloat16)
            torch.cuda.synchronize()
            lib.laff_debug_fuse_profile(buf, 1)
            v = list(buf)
            nf = max(1, v[7])
            units = nf // max(1, len(ks))
            print("%s | cta_group::%d | per projected feature: %s | per unit: tiled %d, normalise+store %d cycles" % (
                name[:28], variant, ", ".join("%s %d" % (NAMES[i], v[i] // nf) for i in range(5)), v[5] // max(1, units), v[6] // max(1, units)), flush=True)
        ops.set_fuse_variant(0)


if __name__ == "__main__":
    main()
