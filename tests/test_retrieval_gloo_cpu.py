"""Host-side logic of the gallery-sharded search on CPU: shard arithmetic and the three collectives over gloo with
world_size 2 and 3, with a numpy stand-in for the CUDA kernels (built from the oracle — test infrastructure only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from laff_b200 import synth
from laff_b200.retrieval import GalleryIndex, Retriever, shard_bounds
from oracle import laff_oracle as O


class NumpyBackend:
    """CPU stand-in with the same contract as laff_b200.retrieval.CudaBackend."""

    def gt_scores(self, q16, g16, gt_local):
        q, g, gl = q16.float().numpy().astype(np.float64), g16.float().numpy().astype(np.float64), gt_local.numpy()
        s = np.where(gl >= 0, np.einsum("ij,ij->i", q, g[np.maximum(gl, 0)]), 0.0)
        return torch.from_numpy(s.astype(np.float32))

    def rank_topk(self, q16, g16, sgt_raw, gt_global, k, scale, col_offset, workspace=None):
        s = (q16.float().numpy().astype(np.float64) @ g16.float().numpy().astype(np.float64).T).astype(np.float32)
        sg = sgt_raw.numpy()[:, None]
        cols = np.arange(s.shape[1])[None, :] + col_offset
        gt = gt_global.numpy()[:, None]
        beats = ((s > sg) | ((s == sg) & (cols > gt))) & (cols != gt)
        count = torch.from_numpy(beats.sum(1).astype(np.int32))
        order = np.lexsort((-cols.repeat(s.shape[0], 0), -s), axis=1)[:, :k]  # score desc, index desc
        tv = np.take_along_axis(s, order, 1) * scale
        ti = order + col_offset
        if tv.shape[1] < k:
            pad = k - tv.shape[1]
            tv = np.concatenate([tv, np.full((s.shape[0], pad), -np.inf, np.float32)], 1)
            ti = np.concatenate([ti, np.full((s.shape[0], pad), -1)], 1)
        return count, torch.from_numpy(tv.astype(np.float32)), torch.from_numpy(ti.astype(np.int32))

    def merge(self, vals, idx, k):
        v = vals.numpy().transpose(1, 0, 2).reshape(vals.shape[1], -1)
        i = idx.numpy().transpose(1, 0, 2).reshape(vals.shape[1], -1)
        order = np.lexsort((-i, -v), axis=1)[:, :k]
        return torch.from_numpy(np.take_along_axis(v, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))

    def dense_topk(self, q16, g16, k, scale, col_offset):
        s = (q16.float().numpy().astype(np.float64) @ g16.float().numpy().astype(np.float64).T).astype(np.float32) * np.float32(scale)
        v, i = O.tie_rule_topk(s, k)
        pad = k - v.shape[1]
        v = np.pad(v, ((0, 0), (0, pad)), constant_values=-np.inf)
        i = np.pad(i + col_offset, ((0, 0), (0, pad)), constant_values=-1)
        return torch.from_numpy(v.astype(np.float32)), torch.from_numpy(i.astype(np.int32))

    def merge_lists(self, vals, idx, k):
        v, i = vals.numpy(), idx.numpy().astype(np.int64)
        order = np.lexsort((-i, -v), axis=1)[:, :k]  # empty slots (-inf, -1) sort last
        return torch.from_numpy(np.take_along_axis(v, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1).astype(np.int32))

    def metrics(self, rank0):
        m = O.metrics_from_rank0(rank0.numpy())
        return torch.tensor(list(m) + [m[5], float(len(rank0))], dtype=torch.float64)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(Q=48, V=301, H=4, dh=16):
    q, g, gt = synth.retrieval_embeddings(77, Q, V, H, dh, sigma=1.5)
    g[13] = g[gt[5]]  # exact tie with a ground truth, across shard boundaries for W >= 2
    g[V - 1] = g[gt[7]]
    return synth.bf16_round(q), synth.bf16_round(g), gt, H


class _StubTxtNet:
    """Stands in for MultiScaleTxtEncoderAttention.encode on CPU: a fixed per-row map to [rows, H, d_h] 16-bit."""

    def encode(self, feats, out16_dtype=None, **kw):
        x = feats["x"].float()
        H = 4
        y = torch.tanh(x * 3.0 + 0.25).reshape(x.shape[0], H, -1)
        return y, y.to(out16_dtype)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q, g, gt, H = _problem()
        lo, hi = shard_bounds(g.shape[0], world, rank)
        idx = GalleryIndex(torch.from_numpy(g[lo:hi]).to(torch.bfloat16), g.shape[0], H, rank, world, backend=NumpyBackend())
        res = idx.search(torch.from_numpy(q).to(torch.bfloat16), torch.from_numpy(gt), k=7)
        lv, li = idx.ranked_lists(torch.from_numpy(q).to(torch.bfloat16), k=150, query_chunk=20)
        # lists + exact ranks in one call (one sweep on the CUDA backend; a rank-only sweep per chunk on backends without it)
        lv3, li3, r3 = idx.ranked_lists(torch.from_numpy(q).to(torch.bfloat16), k=150, query_chunk=20, gt_global=torch.from_numpy(gt))
        lists_ok = torch.equal(r3, res.rank0) and torch.equal(li3, li) and torch.equal(lv3, lv)
        # query fusion is sharded too: every rank encodes 1/W of the queries and the slices are all-gathered; a host
        # caller may hand over only its own slice (total=Q) -- both must reproduce the unsharded embeddings, also when
        # the last ranks get short or empty slices (Q = 5, W = 3)
        retr = Retriever(_StubTxtNet(), idx)
        enc_ok = bool(lists_ok)
        for Qs in (q.shape[0], 5, 1):
            feats = {"x": torch.from_numpy(q[:Qs]).clone()}
            want = _StubTxtNet().encode(feats, out16_dtype=torch.bfloat16)[1].reshape(Qs, -1)
            a, b, _ = retr.query_slice(Qs)
            full = retr.encode_queries(feats)
            pre = retr.encode_queries({"x": feats["x"][a:b]}, total=Qs)
            enc_ok = enc_ok and torch.equal(full, want) and torch.equal(pre, want)
        # the pipelined path (pieces of the batch through pre / sweep / post stages with their own communicators) must
        # give the serial path's answer, for piece counts that divide the batch unevenly and for one piece
        feats = {"x": torch.from_numpy(q).clone()}
        serial = retr.rank(feats, torch.from_numpy(gt), 7)
        for pieces in (1, 3, 5):
            pend = retr.submit(feats, torch.from_numpy(gt), 7, pieces=pieces)
            got, host = pend.result(), pend.to_host()
            enc_ok = enc_ok and all(torch.equal(a, b) for a, b in ((got.rank0, serial.rank0), (got.topk_idx, serial.topk_idx),
                                                                   (got.topk_val, serial.topk_val), (got.metrics, serial.metrics),
                                                                   (host.rank0, serial.rank0)))
        oks = [None] * world
        dist.all_gather_object(oks, bool(enc_ok))
        if rank == 0:
            torch.save({"rank0": res.rank0, "tv": res.topk_val, "ti": res.topk_idx, "m": res.metrics, "lv": lv, "li": li,
                        "enc_ok": all(oks)}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_equals_single_shard(world, tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    q, g, gt, H = _problem()
    single = GalleryIndex(torch.from_numpy(g).to(torch.bfloat16), g.shape[0], H, backend=NumpyBackend())
    ref = single.search(torch.from_numpy(q).to(torch.bfloat16), torch.from_numpy(gt), k=7)
    assert got["enc_ok"]
    assert torch.equal(got["rank0"], ref.rank0)
    assert torch.equal(got["ti"], ref.topk_idx)
    assert torch.allclose(got["tv"], ref.topk_val, atol=0, rtol=0)
    assert torch.equal(got["m"], ref.metrics)
    # writer lists (top-150 of a 301-video gallery, shards of 101..151 videos): sharded == single shard == oracle
    sv, si = single.ranked_lists(torch.from_numpy(q).to(torch.bfloat16), k=150, query_chunk=48)
    assert torch.equal(got["li"], si) and torch.equal(got["lv"], sv)
    np.testing.assert_array_equal(si.numpy(), O.tie_rule_topk((O.mm_mean_heads(q, g, H)).astype(np.float32), 150)[1])
    # and the single-shard answer is the oracle's tie-rule answer on the same operands
    s = (O.mm_mean_heads(q, g, H)).astype(np.float32)
    np.testing.assert_array_equal(ref.rank0.numpy(), O.tie_rule_rank(s, gt))
    np.testing.assert_array_equal(ref.topk_idx.numpy(), O.tie_rule_topk(s, 7)[1])


def test_shard_bounds_cover_the_gallery():
    for V in (0, 1, 7, 8, 1000000):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(V, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == V
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert all(lo <= hi for lo, hi in spans)


def test_weighted_shard_bounds():
    """Shards proportional to per-rank weights: cover the gallery, cut at whole column tiles, equal weights ~ equal
    shards, a slower rank gets fewer rows."""
    V = 1000000
    for W in (2, 4, 8):
        spans = [shard_bounds(V, W, r, [1.0] * W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == V and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % 256 == 0 for lo, _ in spans) and max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 512
    w = [1.0, 0.9, 1.0, 1.1]
    sizes = [b - a for a, b in (shard_bounds(V, 4, r, w) for r in range(4))]
    assert sum(sizes) == V and sizes[1] < sizes[0] < sizes[3] and abs(sizes[1] / V - 0.9 / 4.0) < 1e-3
    assert shard_bounds(100, 3, 1, [1, 1, 1], align=1) == (33, 67)
    with pytest.raises(ValueError):
        shard_bounds(V, 4, 0, [1.0, 1.0])


def test_gallery_index_rejects_wrong_shard():
    with pytest.raises(ValueError):
        GalleryIndex(torch.zeros(5, 8), 100, 2, rank=0, world_size=2, backend=NumpyBackend())


def test_collect_plan_keeps_its_statistical_margins():
    """The threshold path of the writer lists takes the r-th best score of an n_s-row sample as its threshold: the number
    of survivors in the whole shard then has mean r V / n_s and relative spread 1 / sqrt(r).  Whenever a plan is
    returned, mean - 6 sigma must still cover k and mean + 6 sigma must fit the candidate slots; small shards and lists
    too long for the slots must get no plan (dense path)."""
    from laff_b200.retrieval import CudaBackend as B
    assert B.collect_plan(100000, 500) is None and B.collect_plan(8 * B.collect_sample - 1, 10) is None
    assert B.collect_plan(1000000, 0) is None
    seen = 0
    for V in (8 * B.collect_sample, 1000000, 3000000, 20000000):
        for k in (1, 10, 100, 500, 1000, 2000, 2048):
            plan = B.collect_plan(V, k)
            if plan is None:
                continue
            n_s, r = plan
            seen += 1
            mean, rel = r * V / n_s, 6.0 / r ** 0.5
            assert n_s == B.collect_sample and 1 <= r <= 2048
            assert mean * (1 - rel) >= k and mean * (1 + rel) <= B.collect_cap
    assert seen >= 10 and B.collect_plan(1000000, 2000) is not None and B.collect_plan(1000000, 500) is not None


@pytest.mark.parametrize("Q,k", [(3, 1), (7, 5), (1, 10), (8, 4)])
def test_search_result_unpack_any_parity_of_q(Q, k):
    """SearchResult.to_host() packs [metrics (float64) | rank0 | score bits | indices] into one int32 buffer; the float64
    view must stay 8-byte aligned for odd Q (a packed buffer with the metrics last failed there)."""
    from laff_b200.retrieval import SearchResult
    m = torch.arange(8, dtype=torch.float64) * 1.5
    r = torch.arange(Q, dtype=torch.int32)
    v = torch.rand(Q, k)
    i = torch.randint(0, 100, (Q, k), dtype=torch.int32)
    flat = torch.cat([m.view(torch.int32), r, v.reshape(-1).view(torch.int32), i.reshape(-1)])
    h = SearchResult._unpack(flat, Q, (Q, k), 8)
    assert torch.equal(h.metrics, m) and torch.equal(h.rank0, r) and torch.equal(h.topk_val, v) and torch.equal(h.topk_idx, i)
