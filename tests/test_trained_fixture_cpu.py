"""T2 end-to-end parity fixture on the CPU (SURVEY §8c "trained fixture"): tests/golden/trained_laff.npz holds a
checkpoint the UNMODIFIED reference trained for 300 steps (tests/golden/make_golden_trained.py) and the reference's own
ranks / top-10 / R@K / MedR on two held-out collections.  Here the oracle restatement is pinned against it: same
ranks and lists wherever the reference's own fp32 scores resolve them, same metrics."""
import os

import numpy as np
import pytest

from laff_b200 import synth
from oracle import laff_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
NOISE_WINDOW = 2e-6     # two fp32 pipelines (torch / numpy BLAS) agree on a score to ~1e-6 at D = 2048


def load_trained():
    g = np.load(os.path.join(HERE, "golden", "trained_laff.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    vn, cn, tn = [float(x) for x in g["noise"]]
    return g, sd, dict(vis_noise=vn, cap_noise=cn, txt_noise=tn)


def split_state(sd):
    return ({k[len("vis_net."):]: v for k, v in sd.items() if k.startswith("vis_net.")},
            {k[len("txt_net."):]: v for k, v in sd.items() if k.startswith("txt_net.")})


def test_fixture_is_a_trained_model():
    g, sd, _ = load_trained()
    assert g["losses"][0] > 1.5 * g["losses"][-1]                 # the reference's loss went down
    for tag in ("c1", "c2"):
        r1, r5, r10, medr = g[tag + "/metrics"][:4]
        assert 30 < r1 < 90 and r10 > r1 and medr >= 1            # neither chance level nor saturated
        n = int(g[tag + "/n"])
        assert (g[tag + "/rank0"] == 0).mean() * 100 == pytest.approx(r1, abs=1e-9) and g[tag + "/rank0"].shape == (n,)


@pytest.mark.parametrize("tag", ["c1", "c2"])
def test_oracle_reproduces_reference_ranks_on_trained_checkpoint(tag):
    g, sd, noise = load_trained()
    n, H = int(g[tag + "/n"]), int(g["meta"][1])
    vis, txt = synth.latent_collection(int(g[tag + "/seed"]), n, **noise)
    vsd, tsd = split_state(sd)
    ov, _ = O.vis_net_forward(vis, vsd, [synth.VIS_CLIP_FT], H)
    ot, _ = O.txt_net_forward(txt, tsd, ["CLIP_encoder"], H)
    assert np.abs(ot[:8] - g[tag + "/emb_txt_sample"]).max() <= 2e-6 and np.abs(ov[:8] - g[tag + "/emb_vis_sample"]).max() <= 2e-6
    s = O.txt2vis_matrix(ot, ov)
    gt = np.arange(n)
    assert np.abs(s[gt, gt] - g[tag + "/s_gt"]).max() <= NOISE_WINDOW
    rank0 = O.tie_rule_rank(s, gt)
    clean = g[tag + "/gt_gap"] > 2 * NOISE_WINDOW
    assert clean.mean() > 0.98
    np.testing.assert_array_equal(rank0[clean], g[tag + "/rank0"][clean])
    assert np.abs(rank0[~clean] - g[tag + "/rank0"][~clean]).max(initial=0) <= 2
    m = O.metrics_from_rank0(g[tag + "/rank0"])
    np.testing.assert_allclose(m[:4], g[tag + "/metrics"][:4], atol=1e-9)          # metric restatement on the reference's ranks
    mo = O.metrics_from_rank0(rank0)
    assert mo[3] == g[tag + "/metrics"][3] and all(abs(mo[i] - g[tag + "/metrics"][i]) <= 100.0 * (~clean).sum() / n + 1e-9 for i in range(3))
    tv, ti = O.tie_rule_topk(s, 10)
    lists_clean = g[tag + "/min_gap_top11"] > 2 * NOISE_WINDOW
    assert lists_clean.mean() > 0.95
    np.testing.assert_array_equal(ti[lists_clean], g[tag + "/top10"][lists_clean])
    assert np.abs(tv - g[tag + "/top_scores"][:, :10]).max() <= NOISE_WINDOW
