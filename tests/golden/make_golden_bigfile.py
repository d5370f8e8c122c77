"""Golden fixture for the feature-I/O row: a tiny feature directory in the reference's format and what
the UNMODIFIED reference `bigfile.BigFile` returns for a few requests.

    python tests/golden/make_golden_bigfile.py     # writes tests/golden/bigfile/{shape.txt,id.txt,feature.bin,golden.json}
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

REQUESTS = {
    "some": ["vid7", "vid2", "vid2", "nosuch", "vid11"],
    "one": ["vid0"],
    "none": ["nosuch"],
    "by_index": [5, 1, 9],
}


def main():
    mg.install_shims()
    from bigfile import BigFile as RefBigFile
    d = os.path.join(HERE, "bigfile")
    os.makedirs(d, exist_ok=True)
    rng = np.random.RandomState(7)
    n, dims = 13, 6
    names = ["vid%d" % i for i in rng.permutation(n)]
    feats = rng.standard_normal((n, dims)).astype(np.float32)
    open(os.path.join(d, "shape.txt"), "w").write("%d %d" % (n, dims))
    open(os.path.join(d, "id.txt"), "w").write(" ".join(names))  # the space-separated flavour (bigfile.py:19-20)
    feats.tofile(os.path.join(d, "feature.bin"))
    ref = RefBigFile(d)
    out = {"shape": ref.shape(), "names": ref.names}
    for k, req in REQUESTS.items():
        nm, vec = ref.read(req, isname=(k != "by_index"))
        out[k] = {"request": req, "names": nm, "vectors": vec}
    out["read_one"] = {"name": "vid3", "vector": ref.read_one("vid3")}
    nm, vec = ref.readall()
    out["readall"] = {"names": nm, "vectors": vec}
    json.dump(out, open(os.path.join(d, "golden.json"), "w"))
    print("written", d)


if __name__ == "__main__":
    main()
