"""Golden vectors for DualSoftmaxLoss (loss.py:291-310) from the UNMODIFIED reference, with autograd gradients.

    python tests/golden/make_golden_dsl.py     # writes tests/golden/dsl.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
from laff_b200 import synth  # noqa: E402


def main():
    import torch
    mm, rloss, reval, ratt = mg.import_reference()
    crit = rloss.DualSoftmaxLoss()
    out = {}
    for tag, B, H, dh, corr in (("small", 12, 4, 32, 0.0), ("b128", 128, 2, 64, 0.5)):
        r = synth.rng_for(61, "dsl/" + tag)
        vis = r.standard_normal((B, H, dh)).astype(np.float32)
        txt = (vis * corr + r.standard_normal((B, H, dh))).astype(np.float32)
        t, v = torch.from_numpy(txt).requires_grad_(True), torch.from_numpy(vis).requires_grad_(True)
        loss = 0
        per_head = []
        for h in range(H):                                  # model/model.py:2036-2038
            lh = crit(t[:, h, :], v[:, h, :])
            per_head.append(float(lh))
            loss = loss + lh
        loss.backward()
        out.update({tag + "/txt": txt, tag + "/vis": vis, tag + "/loss": np.float64(loss.item()), tag + "/per_head": np.array(per_head),
                    tag + "/d_txt": t.grad.numpy().copy(), tag + "/d_vis": v.grad.numpy().copy()})
        for temp in (1.0, 0.05):                            # other temperatures, first head only
            t2, v2 = torch.from_numpy(txt[:, 0]).requires_grad_(True), torch.from_numpy(vis[:, 0]).requires_grad_(True)
            l2 = crit(t2, v2, temp=temp)
            l2.backward()
            out.update({"%s/temp%g/loss" % (tag, temp): np.float64(l2.item()), "%s/temp%g/d_txt" % (tag, temp): t2.grad.numpy().copy(),
                        "%s/temp%g/d_vis" % (tag, temp): v2.grad.numpy().copy()})
        print(tag, "loss", loss.item())
    np.savez_compressed(os.path.join(HERE, "dsl.npz"), **out)


if __name__ == "__main__":
    main()
