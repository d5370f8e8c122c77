"""bench.py contract pieces that need no GPU: the reference arm (the unmodified reference staged under baseline/_ref when it
is there, else the oracle port of its CPU path) prints
exactly one JSON line with the keys the driver reads, rank != 0 of a multi-rank launch stays silent, and the product
arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--videos", "3000", "--cpu-sample", "8", "--steps", "2", "--warmup", "1", "--gpus", "1"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "queries/sec ranked vs 1M-video gallery" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["value"] > 0 and d["ms_per_step"] > 0
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "laff_reference", "MANIFEST.json"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["gallery"] == 3000 and "workload" in d["config"]


def test_reference_arm_falls_back_to_the_port_without_a_staged_reference(tmp_path, monkeypatch):
    """The port (oracle restatement) is the documented fallback and must give the same ranks as the staged reference's
    own code on the same sample."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from laff_b200 import synth
    from oracle import laff_oracle as O
    ref = bench.staged_reference()
    V, S = 3000, 8
    g = bench.unit_rows(V, torch.Generator().manual_seed(1), "cpu").numpy()
    gt = (np.arange(32) * 97) % V
    q = synth.unit_heads(g[gt] + 6.0 * bench.unit_rows(32, torch.Generator().manual_seed(2), "cpu").numpy(), 8)
    out = O.retrieve_cpu(q[:S], g, gt[:S], 8, k=10, chunk=100000, threads=2)
    if ref is None:
        return
    _, m_ref = bench.reference_cpu_steps(ref, q, g, gt, S, 1, 0, 2)
    port_metrics = out[-1]
    assert [float(x) for x in m_ref[:4]] == [float(x) for x in port_metrics[:4]], (m_ref, port_metrics)


def test_reference_arm_other_ranks_do_no_work():
    r = _run(["--impl", "reference", "--videos", "3000", "--cpu-sample", "8", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--queries", "16", "--videos", "64", "--steps", "1", "--warmup", "1", "--mode-b-steps", "0"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
