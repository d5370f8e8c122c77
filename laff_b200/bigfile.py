"""Feature I/O of the reference (`bigfile.py:13-234`) plus the bulk path the retrieval stack needs.

A feature directory holds `shape.txt` ("<rows> <dims>"), `id.txt` (names separated by newlines or, in older dumps, by
spaces) and `feature.bin` (row-major float32, no header).  The reference opens the file and seeks once per requested
vector and converts every vector to a Python list (`bigfile.py:211-234`); here the file is memory-mapped, the name
API (`read`, `read_one`, `readall`, `shape`) keeps the reference's contract, and `rows()` / `to_device()` move whole
shards: contiguous row ranges are read straight into pinned staging buffers and copied to the GPU asynchronously,
double-buffered, so a gallery shard streams at disk / PCIe speed without per-item Python work.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch


class BigFile:
    def __init__(self, datadir: str, bin_file: str = "feature.bin"):
        with open(os.path.join(datadir, "shape.txt")) as f:
            self.nr_of_images, self.ndims = list(map(int, f.readline().split()))
        id_file = os.path.join(datadir, "id.txt")
        with open(id_file, "r") as f:
            text = f.read().strip()
        self.names = text.split("\n")
        if len(self.names) != self.nr_of_images:  # older dumps separate the names by spaces (bigfile.py:19-20)
            self.names = text.split(" ")
        assert len(self.names) == self.nr_of_images
        self.name2index = dict(zip(self.names, range(self.nr_of_images)))
        self.binary_file = os.path.join(datadir, bin_file)
        expect = self.nr_of_images * self.ndims * 4
        have = os.path.getsize(self.binary_file)
        if have < expect:
            raise IOError("%s holds %d bytes, shape.txt promises %d" % (self.binary_file, have, expect))
        self._mm: Optional[np.memmap] = None
        print("[%s] %dx%d instances loaded from %s" % (self.__class__.__name__, self.nr_of_images, self.ndims, datadir))

    # ------------------------------------------------------------------ bulk access
    def matrix(self) -> np.memmap:
        """The whole file as a read-only [rows, dims] float32 memory map."""
        if self._mm is None:
            self._mm = np.memmap(self.binary_file, dtype=np.float32, mode="r", shape=(self.nr_of_images, self.ndims))
        return self._mm

    def rows(self, lo: int, hi: int) -> np.ndarray:
        """Rows [lo, hi) as a float32 array (one contiguous read)."""
        if not 0 <= lo <= hi <= self.nr_of_images:
            raise IndexError("rows [%d, %d) outside [0, %d)" % (lo, hi, self.nr_of_images))
        return np.array(self.matrix()[lo:hi])

    def indices(self, names: Sequence[str]) -> np.ndarray:
        """Row index of every name, in the given order (KeyError for unknown names)."""
        return np.fromiter((self.name2index[n] for n in names), dtype=np.int64, count=len(names))

    def to_device(self, lo: int, hi: int, device, chunk_rows: int = 65536, out: Optional[torch.Tensor] = None,
                  stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """Rows [lo, hi) -> CUDA float32 tensor [hi - lo, dims].  The file is read in `chunk_rows` pieces directly into
        two pinned staging buffers (`readinto`, no intermediate numpy copy) whose host->device copies are asynchronous:
        reading piece i + 1 from disk overlaps the DMA of piece i."""
        if not 0 <= lo <= hi <= self.nr_of_images:
            raise IndexError("rows [%d, %d) outside [0, %d)" % (lo, hi, self.nr_of_images))
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("to_device streams to a CUDA device; use rows() for host arrays")
        n = hi - lo
        if out is None:
            out = torch.empty((n, self.ndims), dtype=torch.float32, device=device)
        elif tuple(out.shape) != (n, self.ndims) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 [%d, %d] tensor" % (n, self.ndims))
        if n == 0:
            return out
        chunk_rows = max(1, min(chunk_rows, n))
        stage = [torch.empty((chunk_rows, self.ndims), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [None, None]
        st = stream or torch.cuda.current_stream(device)
        with open(self.binary_file, "rb", buffering=0) as f:
            f.seek(lo * self.ndims * 4)
            for i, s in enumerate(range(0, n, chunk_rows)):
                e = min(n, s + chunk_rows)
                buf = stage[i & 1]
                if done[i & 1] is not None:
                    done[i & 1].synchronize()  # the copy that last used this staging buffer has finished
                view = buf[: e - s].numpy().reshape(-1).view(np.uint8)
                got = 0
                while got < view.size:  # raw files may return short reads
                    r = f.readinto(memoryview(view)[got:])
                    if not r:
                        raise IOError("unexpected end of %s" % self.binary_file)
                    got += r
                with torch.cuda.stream(st):
                    out[s:e].copy_(buf[: e - s], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(st)
                done[i & 1] = ev
        for ev in done:
            if ev is not None:
                ev.synchronize()
        return out

    # ------------------------------------------------------------------ the reference's name API
    def _select(self, requested: Iterable, isname: bool) -> List[Tuple[int, str]]:
        requested = set(requested)
        if isname:
            pairs = [(self.name2index[x], x) for x in requested if x in self.name2index]
        else:
            assert min(requested) >= 0
            assert max(requested) < len(self.names)
            pairs = [(x, self.names[x]) for x in requested]
        pairs.sort(key=lambda v: v[0])
        return pairs

    def read_array(self, requested: Iterable, isname: bool = True) -> Tuple[List[str], np.ndarray]:
        """`read` without the list conversion: (names sorted by file position, float32 [n, dims])."""
        pairs = self._select(requested, isname)
        if not pairs:
            return [], np.zeros((0, self.ndims), dtype=np.float32)
        idx = np.fromiter((p[0] for p in pairs), dtype=np.int64, count=len(pairs))
        return [p[1] for p in pairs], np.array(self.matrix()[idx])

    def read(self, requested: Iterable, isname: bool = True):
        """bigfile.py:197-234: duplicates collapse, unknown names are dropped silently, the result is ordered by
        position in the file; vectors come back as Python lists."""
        names, arr = self.read_array(requested, isname)
        return names, [row.tolist() for row in arr]

    def readall(self, isname: bool = True):
        """bigfile.py:71-96."""
        return self.read(self.names)

    def read_one(self, name):
        """bigfile.py:222-226 (IndexError for an unknown name, like `vectors[0]` on the empty result)."""
        return self.read([name])[1][0]

    def shape(self):
        return [self.nr_of_images, self.ndims]


def load_features(bigfiles, ids: Sequence[str], device, chunk_rows: int = 65536):
    """{feature name: BigFile} + the ids of a gallery shard -> {feature name: CUDA float32 [len(ids), dims]}, row i
    holding the feature of ids[i] (what `VisionDataset.__getitem__` + `collate_vision` assemble item by item,
    data_provider.py:38-73, :380-498).  Every feature file may store the videos in its own order: a shard that is a
    contiguous ascending run of a file streams through `BigFile.to_device`; otherwise rows are gathered from the
    memory map in file order (sequential-friendly) through a pinned buffer and un-permuted on the device."""
    device = torch.device(device)
    out = {}
    for name, bf in bigfiles.items():
        idx = bf.indices(ids)
        n = len(idx)
        if n and np.array_equal(idx, np.arange(idx[0], idx[0] + n)):
            out[name] = bf.to_device(int(idx[0]), int(idx[0]) + n, device, chunk_rows)
            continue
        order = np.argsort(idx, kind="stable")
        dst = torch.empty((n, bf.ndims), dtype=torch.float32, device=device)
        mm = bf.matrix()
        stage = torch.empty((min(chunk_rows, max(n, 1)), bf.ndims), dtype=torch.float32).pin_memory()
        for s in range(0, n, chunk_rows):
            e = min(n, s + chunk_rows)
            np.take(mm, idx[order[s:e]], axis=0, out=stage[: e - s].numpy())
            dst[s:e].copy_(stage[: e - s], non_blocking=True)
            torch.cuda.current_stream(device).synchronize()  # the staging buffer is reused
        inv = torch.from_numpy(order).to(device)
        res = torch.empty_like(dst)
        res[inv] = dst
        out[name] = res
    return out


def write_bigfile(datadir: str, names: Sequence[str], features: np.ndarray, bin_file: str = "feature.bin") -> None:
    """Write a feature directory in the reference's format (the inverse of BigFile; used to build synthetic collections)."""
    features = np.ascontiguousarray(features, dtype=np.float32)
    if features.ndim != 2 or features.shape[0] != len(names):
        raise ValueError("features must be [len(names), dims]")
    if any((" " in n) or ("\n" in n) for n in names):
        raise ValueError("names must not contain spaces or newlines")
    os.makedirs(datadir, exist_ok=True)
    with open(os.path.join(datadir, "shape.txt"), "w") as f:
        f.write("%d %d" % features.shape)
    with open(os.path.join(datadir, "id.txt"), "w") as f:
        f.write("\n".join(names))
    features.tofile(os.path.join(datadir, bin_file))
