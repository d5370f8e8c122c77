"""Feature I/O on the GPU box: streaming a BigFile shard to the device, and the on-disk collection ->
gallery index -> ranking path giving the same embeddings / ranks as the in-memory path."""
import numpy as np
import pytest
import torch

from laff_b200 import config as cfg
from laff_b200 import model as M
from laff_b200 import synth
from laff_b200.bigfile import BigFile, load_features, write_bigfile
from laff_b200.retrieval import GalleryIndex, Retriever
from helpers import load_numpy_state

pytestmark = pytest.mark.gpu


def test_to_device_streams_exact_rows(tmp_path):
    rng = np.random.RandomState(1)
    feats = rng.standard_normal((70001, 48)).astype(np.float32)
    write_bigfile(str(tmp_path / "f"), ["v%d" % i for i in range(len(feats))], feats)
    bf = BigFile(str(tmp_path / "f"))
    for lo, hi, chunk in ((0, 70001, 8192), (5, 69999, 65536), (123, 124, 7), (40, 40, 16)):
        t = bf.to_device(lo, hi, "cuda", chunk_rows=chunk)
        assert t.is_cuda and np.array_equal(t.cpu().numpy(), feats[lo:hi])
    with pytest.raises(IndexError):
        bf.to_device(0, 80000, "cuda")


def test_collection_on_disk_equals_in_memory(tmp_path):
    c = cfg.laff_config(4096, 8, synth.DIMS)
    dims = dict(c.vis_fc_layers[0])
    V, Q = 700, 90
    rng = np.random.RandomState(2)
    vis_ids = ["video%04d" % i for i in rng.permutation(V)]
    feats, files = {}, {}
    for name, d in dims.items():
        x = rng.standard_normal((V, d)).astype(np.float32)
        feats[name] = x                                           # row i belongs to vis_ids[i]
        order = np.arange(V) if name == synth.VIS_CLIP_FT else rng.permutation(V)  # every file has its own row order
        write_bigfile(str(tmp_path / name), [vis_ids[j] for j in order], x[order])
        files[name] = BigFile(str(tmp_path / name))
    vis_net = M.VisMutiTransformNetAddAttnetion(c, c.vis_fc_layers[0])
    load_numpy_state(vis_net, {k: synth.param(3, k, tuple(v.shape)) for k, v in vis_net.state_dict().items()})
    vis_net = vis_net.cuda().eval()
    txt_net = M.MultiScaleTxtEncoderAttention(c)
    load_numpy_state(txt_net, {k: synth.param(4, k, tuple(v.shape)) for k, v in txt_net.state_dict().items()})
    txt_net = txt_net.cuda().eval()
    loaded = load_features(files, vis_ids, "cuda", chunk_rows=256)
    for name in dims:
        assert np.array_equal(loaded[name].cpu().numpy(), feats[name]), name
    a = GalleryIndex.from_features(vis_net, loaded, V)
    b = GalleryIndex.from_features(vis_net, {k: torch.from_numpy(v) for k, v in feats.items()}, V)
    assert torch.equal(a.g16, b.g16)
    q = {"gru": torch.randn(Q, 1024), "bow": torch.zeros(Q, synth.DIMS["bow"]), "w2v": torch.randn(Q, 500), "clip": torch.randn(Q, 512)}
    gt = torch.arange(Q, dtype=torch.int32)
    ra = Retriever(txt_net, a).rank(q, gt, 10)
    rb = Retriever(txt_net, b).rank(q, gt, 10)
    assert torch.equal(ra.rank0, rb.rank0) and torch.equal(ra.topk_idx, rb.topk_idx)
