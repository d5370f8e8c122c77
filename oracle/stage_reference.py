"""Stages the UNMODIFIED pure-Python reference (ruc-aimc-lab/LAFF) for the CPU arm of bench.py -- TEST / BASELINE
INFRASTRUCTURE, never imported by the product path.

The reference has no compiled code and no setup.py, so "installing" it is a file copy: the modules its
evaluation path imports (model/, loss.py, evaluation.py, util.py, ...) are copied verbatim from /root/reference into
`baseline/_ref/laff_reference/` -- the location the bench contract reserves for the unmodified reference -- together
with a MANIFEST of their sha256 sums.  `baseline/_ref/` is git-ignored (no reference source enters the history) but not
gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not exist, and `bench.py --impl
reference` can time the reference's own functions there (`kind: "reference"`).
Run by `__graft_entry__.build()` whenever /root/reference is present.

    python oracle/stage_reference.py
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "laff_reference")
SOURCE = os.environ.get("LAFF_REFERENCE", "/root/reference")
# what `import model.model`, `import evaluation`, `import loss` pull in (model/model.py:1-26)
FILES = ["__init__.py", "bigfile.py", "common.py", "evaluation.py", "generic_utils.py", "loss.py", "textlib.py", "txt2vec.py",
         "util.py", "stopwords_en.txt", "stopwords_zh.txt", "model/Attention.py", "model/ReRank.py", "model/model.py", "model/clip/__init__.py", "model/clip/clip.py",
         "model/clip/model.py", "model/clip/simple_tokenizer.py", "model/clip/bpe_simple_vocab_16e6.txt.gz",
         "configs/__init__.py", "configs/base_config.py", "configs/laff.py"]


def stage(source: str = SOURCE, dest: str = DEST) -> str | None:
    """Copies the files; returns the destination, or None when the reference tree is not here (GPU box: use what
    travelled with the snapshot)."""
    if not os.path.isdir(source):
        return dest if os.path.exists(os.path.join(dest, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src = os.path.join(source, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):
                os.makedirs(os.path.dirname(os.path.join(dest, rel)), exist_ok=True)
                open(os.path.join(dest, rel), "a").close()
                continue
            raise FileNotFoundError(src)
        dst = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": source, "files": manifest}, open(os.path.join(dest, "MANIFEST.json"), "w"), indent=1)
    return dest


if __name__ == "__main__":
    print(stage())
