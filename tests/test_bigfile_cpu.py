"""Feature I/O: laff_b200.bigfile.BigFile against what the unmodified reference bigfile.BigFile returned
for the committed fixture directory (tests/golden/bigfile, written by tests/golden/make_golden_bigfile.py)."""
import json
import os

import numpy as np
import pytest

from laff_b200.bigfile import BigFile, write_bigfile

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "bigfile")
GOLD = json.load(open(os.path.join(FIX, "golden.json")))


def test_name_api_matches_reference():
    bf = BigFile(FIX)
    assert bf.shape() == GOLD["shape"] and bf.names == GOLD["names"]
    for key in ("some", "one", "none", "by_index"):
        names, vecs = bf.read(GOLD[key]["request"], isname=(key != "by_index"))
        assert names == GOLD[key]["names"] and vecs == GOLD[key]["vectors"], key   # float32 -> Python float: exact
    assert bf.read_one(GOLD["read_one"]["name"]) == GOLD["read_one"]["vector"]
    names, vecs = bf.readall()
    assert names == GOLD["readall"]["names"] and vecs == GOLD["readall"]["vectors"]
    with pytest.raises(IndexError):
        bf.read_one("nosuch")


def test_bulk_rows_and_roundtrip(tmp_path):
    rng = np.random.RandomState(3)
    names = ["v%05d" % i for i in range(1000)]
    feats = rng.standard_normal((1000, 37)).astype(np.float32)
    write_bigfile(str(tmp_path / "feat"), names, feats)
    bf = BigFile(str(tmp_path / "feat"))
    assert bf.shape() == [1000, 37]
    assert np.array_equal(bf.rows(0, 1000), feats) and np.array_equal(bf.rows(123, 457), feats[123:457])
    assert bf.rows(5, 5).shape == (0, 37)
    nm, arr = bf.read_array(["v00999", "v00000", "v00500", "v00500"])
    assert nm == ["v00000", "v00500", "v00999"] and np.array_equal(arr, feats[[0, 500, 999]])
    assert np.array_equal(bf.indices(["v00007", "v00003"]), [7, 3])
    with pytest.raises(IndexError):
        bf.rows(10, 2000)
    with pytest.raises(KeyError):
        bf.indices(["nosuch"])


def test_truncated_file_is_rejected(tmp_path):
    d = tmp_path / "bad"
    write_bigfile(str(d), ["a", "b"], np.zeros((2, 4), np.float32))
    with open(d / "feature.bin", "r+b") as f:
        f.truncate(20)
    with pytest.raises(IOError):
        BigFile(str(d))
