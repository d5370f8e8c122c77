"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares, the
ctypes binding agrees with the header, and the product path refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import laff_b200
from laff_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "laff_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(laff_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_loads():
    assert os.path.exists(_capi.LIB_PATH), "run `python -m laff_b200.build` (or __graft_entry__.build())"
    lib = _capi.lib()
    assert lib.laff_abi_version() == 1


def test_every_header_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 20
    raw = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "%s declared in include/laff_b200.h but not exported" % n
        assert n in _capi.SIGNATURES, "%s has no ctypes signature in laff_b200/_capi.py" % n
    for n in _capi.SIGNATURES:
        assert n in names, "%s bound in _capi.py but not declared in the header" % n


def test_argument_counts_match_header():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _capi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert n == len(args), "%s: header has %d parameters, ctypes binding %d" % (name, n, len(args))


def test_tuning_roundtrip_and_validation():
    from laff_b200 import ops
    before = ops.get_tuning()
    ops.set_tuning(cta_group=1, chunk_tiles=3, m_group=7)
    assert ops.get_tuning() == (1, 3, 7)
    ops.set_tuning(*before)
    with pytest.raises(laff_b200.LaffError):
        ops.set_tuning(cta_group=5)


def test_workspace_queries_are_pure_host_functions():
    lib = _capi.lib()
    assert lib.laff_sim_gt_workspace_bytes(10000, 4096) >= 10000 * 4096 * 2
    assert lib.laff_sim_rank_workspace_bytes(10000, 1000000, 4096) >= 10000 * (16 * 8 + 4)
    assert lib.laff_sim_rank_workspace_bytes(0, 5, 8) == 0
    assert lib.laff_mrl_workspace_bytes(128, 8, 512) > 2 * 128 * 8 * 512 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    """Without a CUDA device the product path raises; it never routes to the oracle or to torch ops."""
    from laff_b200 import ops, loss, evaluation, model
    x = torch.randn(4, 64)
    with pytest.raises(laff_b200.LaffError):
        ops.l2norm_quantize(x, 1)
    with pytest.raises(laff_b200.LaffError):
        loss.cosine_sim(x, x)
    with pytest.raises(laff_b200.LaffError):
        evaluation.eval_qry2retro(x.numpy()[:, :4])
    with pytest.raises(laff_b200.LaffError):
        model.Attention_1(64, with_ave=False)(torch.randn(2, 3, 64))
    # a compute call that reaches the library reports "no device" through the C ABI error channel
    lib = _capi.lib()
    rc = lib.laff_rank_metrics(None, 0, None, None)
    assert rc != 0 and lib.laff_last_error()


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "laff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("no oracle", ""), "%s mentions the oracle" % f


def test_product_path_never_touches_the_oracle_or_the_staged_reference():
    """oracle/ (and the reference staged under baseline/_ref) is test / baseline infrastructure: nothing under laff_b200/
    may import, open or execute it -- only tests/, __graft_entry__.smoke() and bench.py's CPU legs do."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|baseline[/\\]_ref|oracle[/\\]_ref|ref_loader|laff_oracle", re.M)
    for path in glob.glob(os.path.join(root, "laff_b200", "**", "*.py"), recursive=True) + glob.glob(os.path.join(root, "laff_b200", "csrc", "*")):
        assert not pat.search(open(path, errors="ignore").read()), path
