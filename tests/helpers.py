"""Shared helpers for the parity tests."""
import numpy as np
import torch

from laff_b200 import synth


def sd_from_npz(d, prefix):
    return {k[len(prefix):]: d[k] for k in d.files if k.startswith(prefix)}


def load_numpy_state(module: torch.nn.Module, sd_np, strict=True):
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    missing, unexpected = module.load_state_dict(sd, strict=False)
    if strict:
        assert not missing, "missing keys: %s" % missing
        assert not unexpected, "unexpected keys: %s" % unexpected
    module.eval()
    return module


def small_dims(d):
    """Feature dims of a small golden fusion case."""
    vd = dict(zip([str(n) for n in d["vis_names"]], [int(x) for x in d["vis_dims"]]))
    g, b, w, c = [int(x) for x in d["txt_dims"]]
    return {"clip": c, "gru": g, "bow": b, "w2v": w, "tf": vd[synth.VIS_TF], "x3d": vd[synth.VIS_X3D],
            "ircsn": vd[synth.VIS_IRCSN], "c3d": vd.get(synth.VIS_C3D, 0)}


def cuda(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


def max_abs(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b))) if a.size else 0.0
