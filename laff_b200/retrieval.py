"""Gallery-sharded retrieval: fused similarity + exact rank + top-k + metrics, one process per GPU.

Replaces the reference's evaluation pipeline for large galleries — ``W2VVPP.predict`` building a dense host score
matrix (model/model.py:1018-1079), ``np.argsort`` over every row (predictor.py:232), the per-query ground-truth search
(predictor.py:239-244) and ``evaluation.eval`` (evaluation.py:92-109) — with one tensor-core sweep per gallery shard
that never materialises the Q x V matrix.

Sharding: rank r of W holds gallery rows [r*ceil(V/W), ...) as 16-bit unit-norm embeddings; queries are
replicated.  Exchange steps, all tiny, over torch.distributed (NCCL on GPUs; gloo in the CPU tests):
  1. all_reduce(sum) of s_gt[Q]          (each ground truth is owned by exactly one shard, others contribute 0)
  2. all_reduce(sum) of count[Q] int32   -> exact global rank0
  3. all_gather of the per-shard ordered top-k lists + a k-way merge under the tie rule (global indices).
With one process (W = 1) no collective is issued.

The local compute is delegated to a backend object; the product backend is :class:`CudaBackend` (the sm_100a
kernels).  Tests inject a numpy backend to exercise the sharding / collective logic on CPU with gloo.
"""
from __future__ import annotations

import contextlib
import types
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops


def shard_bounds(V: int, world_size: int, rank: int, weights=None, align: int = 256) -> Tuple[int, int]:
    """Contiguous row range of the gallery held by `rank` (the last shards may be short or empty).

    weights (one positive number per rank, e.g. from calibrate_rank_weights): shard sizes proportional to them instead
    of equal, cut at multiples of `align` rows (whole column tiles of the sweep) -- the GPUs of one box do not sustain
    the same clock under the power cap, and a step is as slow as its slowest shard."""
    if weights is None:
        per = (V + world_size - 1) // world_size
        lo = min(V, rank * per)
        return lo, min(V, lo + per)
    w = [max(float(x), 0.0) for x in weights]
    if len(w) != world_size or sum(w) <= 0:
        raise ValueError("shard_bounds: need %d positive weights, got %r" % (world_size, weights))
    cuts, acc = [0], 0.0
    for x in w[:-1]:
        acc += x / sum(w)
        c = int(round(V * acc / align)) * align
        cuts.append(min(V, max(cuts[-1], c)))
    cuts.append(V)
    return cuts[rank], cuts[rank + 1]


@torch.no_grad()
def calibrate_rank_weights(device, world_size: int, group=None, seconds: float = 1.5, heads: int = 8, dim: int = 4096,
                           clamp: float = 0.15):
    """Relative sweep throughput of every rank's GPU under sustained load, for weighted gallery shards.  Every rank runs
    the similarity sweep (2560 queries x 131 072 synthetic videos, one kernel launch of ~2.7 ms) back to back for
    `seconds` -- long enough for the power cap to settle -- and times the second half; weights = 1 / time, clamped to
    +-`clamp` around their mean (a guard against a noisy measurement), identical on all ranks after one all_gather."""
    if world_size == 1:
        return [1.0]
    Q, V = 2560, 131072
    gen = torch.Generator(device=device).manual_seed(99)
    g = torch.randn(V, dim, generator=gen, device=device).to(torch.float16)
    q = torch.randn(Q, dim, generator=gen, device=device).to(torch.float16)
    gt = torch.zeros(Q, dtype=torch.int32, device=device)
    sgt = torch.full((Q,), 1e30, dtype=torch.float32, device=device)
    ws = None
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    t_end = time.perf_counter() + seconds / 2
    while time.perf_counter() < t_end:                      # warm the clocks into their sustained state
        for _ in range(8):
            ops.sim_rank_topk(q, g, sgt, gt, 10, 1.0 / heads)
        torch.cuda.synchronize(device)
    n = 0
    e0.record()
    t_end = time.perf_counter() + seconds / 2
    while time.perf_counter() < t_end:
        for _ in range(8):
            ops.sim_rank_topk(q, g, sgt, gt, 10, 1.0 / heads)
        n += 8
        torch.cuda.synchronize(device)
    e1.record()
    torch.cuda.synchronize(device)
    mine = torch.tensor([e0.elapsed_time(e1) / max(1, n)], dtype=torch.float64, device=device)
    allt = torch.empty(world_size, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allt, mine, group=group)
    inv = 1.0 / allt.cpu()
    w = inv / inv.mean()
    w = w.clamp(1.0 - clamp, 1.0 + clamp)
    return [float(x) for x in (w / w.sum())]


class CudaBackend:
    """Local shard compute on the B200 through the C ABI (no fallback)."""

    def gt_scores(self, q16, g16, gt_local):
        return ops.sim_gt_scores(q16, g16, gt_local)

    def rank_topk(self, q16, g16, sgt_raw, gt_global, k, scale, col_offset, workspace=None):
        return ops.sim_rank_topk(q16, g16, sgt_raw, gt_global, k, scale=scale, col_offset=col_offset, workspace=workspace)

    def merge(self, vals, idx, k):
        return ops.topk_merge(vals, idx, k)

    def metrics(self, rank0):
        return ops.rank_metrics(rank0)

    dense_rows = 256          # queries per dense chunk: 256 x 1 M fp32 scores = 1 GB
    collect_cap = 8192        # candidate slots per query of the threshold path (sorted whole in shared memory)
    collect_sample = 65536    # gallery rows the threshold is estimated from

    def _dense_topk(self, q16, g16, k, scale, col_offset):
        outs = []
        for lo in range(0, q16.shape[0], self.dense_rows):
            s = ops.sim_dense(q16[lo:lo + self.dense_rows], g16, scale)
            tv, ti = ops.topk_dense(s, k)
            outs.append((tv, torch.where(ti >= 0, ti + col_offset, ti)))
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    @classmethod
    def collect_plan(cls, V: int, k: int):
        """Sample rank whose score serves as the threshold of the collect path, or None when the dense path is the
        better choice.  The threshold is the r-th best score among the first n_s gallery rows; the number of scores
        above it in the whole shard then has mean ~ r V / n_s and relative spread ~ 1 / sqrt(r): r is the smallest rank
        with mean - 6 sigma >= k, and the plan is dropped if mean + 6 sigma would not fit the candidate slots."""
        n_s = cls.collect_sample
        if V < 8 * n_s or k < 1:
            return None
        ratio = V / n_s
        r = max(16, int(k / ratio))
        while r * ratio * (1.0 - 6.0 / r ** 0.5) < k:
            r += max(1, r // 16)
        if r * ratio * (1.0 + 6.0 / r ** 0.5) > cls.collect_cap or r > 2048:
            return None
        return n_s, r

    lists_with_rank = True   # dense_topk(..., sgt, gt) can count the ground truth's rank in the same sweep

    def dense_topk(self, q16, g16, k, scale, col_offset, sgt=None, gt=None):
        """Ranked list of the k best local videos per query (k <= 2048).  Large shards: a threshold just below the k-th
        best is read off a sample of the shard, one sweep keeps the scores above it (laff_sim_collect, ~2k of V per
        query; no Q x V matrix) and the survivors are sorted; if a query ends with fewer than k or more than the slot
        count (a gallery whose first rows are not representative) the chunk is redone densely.  Small shards: dense
        scores of 256-query chunks + laff_topk_dense.  Both give the tie-rule order of the full row.
        With sgt (raw ground-truth scores) and gt (global ground-truth indices) a third result is returned: the local
        count of videos ranked above the ground truth, taken in the same sweep on the threshold path (None on the dense
        path: the caller then counts with a rank-only sweep)."""
        V = g16.shape[0]
        want_rank = sgt is not None
        plan = self.collect_plan(V, k)
        if plan is None:
            out = self._dense_topk(q16, g16, k, scale, col_offset)
            return out + (None,) if want_rank else out
        n_s, r = plan
        sv, _ = ops.topk_dense(ops.sim_dense(q16, g16[:n_s], scale), r)
        thr = sv[:, r - 1].contiguous()
        got = ops.sim_collect(q16, g16, thr, self.collect_cap, scale, col_offset, sgt_raw=sgt, gt_global=gt)
        count, cv, ci = got[:3]
        if bool(((count < k) | (count > self.collect_cap)).any()):       # one host read per chunk
            out = self._dense_topk(q16, g16, k, scale, col_offset)
            return out + (got[3],) if want_rank else out                 # the rank count does not depend on the lists
        out = ops.topk_dense(cv, k, idx_in=ci)
        return out + (got[3],) if want_rank else out

    def merge_lists(self, vals, idx, k):
        """vals/idx [Q, n] candidates carrying global indices (-1 = empty) -> the k best by the tie rule."""
        return ops.topk_dense(vals, k, idx_in=idx)


@dataclass
class SearchResult:
    rank0: torch.Tensor      # int32 [Q]  0-based rank of the ground truth over the whole gallery
    topk_val: torch.Tensor   # fp32 [Q, k] scores (mean over heads of the per-head cosine)
    topk_idx: torch.Tensor   # int32 [Q, k] global gallery indices, ordered by (score desc, index desc)
    metrics: torch.Tensor    # float64 [8]: R@1, R@5, R@10, MedR, MeanR, MIR, mAP, Q

    def to_host(self) -> "SearchResult":
        """The same result in host memory, fetched with ONE device->host copy and one synchronisation (four separate
        `.cpu()` calls cost four round trips -- visible once a step is a few milliseconds, as at 8 GPUs)."""
        if not self.rank0.is_cuda:
            return self
        host, ev = self._pack_to_pinned()
        ev.synchronize()
        return self._unpack(host, self.rank0.numel(), tuple(self.topk_val.shape), self.metrics.numel())

    def _pack_to_pinned(self):
        """Enqueues the packed device->host copy on the current stream; returns (pinned int32 buffer, completion event)."""
        # the float64 metrics go first so that their view starts at an 8-byte aligned offset whatever Q is
        parts = [self.metrics.double().reshape(-1).view(torch.int32), self.rank0.to(torch.int32).reshape(-1),
                 self.topk_val.float().reshape(-1).view(torch.int32), self.topk_idx.to(torch.int32).reshape(-1)]
        flat = torch.cat(parts)
        host = torch.empty(flat.shape, dtype=torch.int32, pin_memory=True)
        host.copy_(flat, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(flat.device))
        return host, ev

    @staticmethod
    def _unpack(host: torch.Tensor, Q: int, klist_shape, n_metrics: int) -> "SearchResult":
        """Inverse of the packing in to_host(): int32 words [metrics (2 words each) | rank0 | score bits | indices]."""
        n_m, n_k = 2 * n_metrics, klist_shape[0] * klist_shape[1]
        o1 = n_m + Q
        return SearchResult(host[n_m:o1], host[o1:o1 + n_k].view(torch.float32).reshape(klist_shape),
                            host[o1 + n_k:o1 + 2 * n_k].reshape(klist_shape), host[:n_m].view(torch.float64))


class PendingSearch:
    """Handle of a search submitted to the pipelined path (Retriever.submit): the result tensors are being produced on the
    pipeline's streams.  `result()` makes the caller's current stream wait for them; `to_host()` returns host copies
    (one packed device->host transfer, enqueued at submit time when fetch=True so that only its completion is awaited)."""

    def __init__(self, res: SearchResult, done, stream, packed=None):
        self._res, self._done, self._stream, self._packed = res, done, stream, packed

    def result(self) -> SearchResult:
        if self._done is not None:
            cur = torch.cuda.current_stream(self._res.rank0.device)
            cur.wait_event(self._done)
            for t in (self._res.rank0, self._res.topk_val, self._res.topk_idx, self._res.metrics):
                t.record_stream(cur)
        return self._res

    def to_host(self) -> SearchResult:
        if self._done is None:
            return self._res.to_host()
        if self._packed is None:
            with torch.cuda.stream(self._stream):
                self._packed = self._res._pack_to_pinned()
        host, ev = self._packed
        ev.synchronize()
        r = self._res
        return SearchResult._unpack(host, r.rank0.numel(), tuple(r.topk_val.shape), r.metrics.numel())


class GalleryIndex:
    """The local shard of a gallery of fused video embeddings, resident in HBM (the reference's record_emb=True cache,
    model/model.py:1026-1034, kept on the device and in the tensor-core operand type)."""

    def __init__(self, shard16: torch.Tensor, total: int, heads: int, rank: int = 0, world_size: int = 1,
                 group=None, backend=None, weights=None):
        lo, hi = shard_bounds(total, world_size, rank, weights)
        if shard16.shape[0] != hi - lo:
            raise ValueError("shard has %d rows, expected %d for rank %d/%d of a %d-row gallery"
                             % (shard16.shape[0], hi - lo, rank, world_size, total))
        self.g16 = shard16
        self.total, self.heads = total, heads
        self.rank, self.world_size, self.group = rank, world_size, group
        self.lo, self.hi = lo, hi
        self.backend = backend or CudaBackend()
        self._ws = None
        self.timers = None  # set to a list to collect (start, end) CUDA events around every sweep launch

    @classmethod
    @torch.no_grad()
    def from_features(cls, vis_net, vis_input, total: int, rank: int = 0, world_size: int = 1, group=None,
                      backend=None, frame_input=None, out16_dtype=None):
        """Fuse this rank's shard of raw video features into the resident 16-bit gallery (the reference's
        `vis_net(vis_input)` loop over the gallery loader, model/model.py:1036-1049, data-parallel over the shard; no
        communication).  vis_input: dict name -> [rows of this shard, d_l] tensors (host or device); frame_input: the
        LAFF-ml frame-feature dict; out16_dtype: torch.float16 / torch.bfloat16 (default: the type of the current
        precision, loss.get_precision()) -- the projection operands follow the same type."""
        from . import loss as _loss
        out16_dtype = out16_dtype or _loss.operand_dtype()
        precision = "fp16" if out16_dtype == torch.float16 else "bf16"
        if frame_input is not None:
            _, g16 = vis_net.encode(vis_input, frame_input, out16_dtype=out16_dtype, precision=precision, want_f32=False)
        else:
            _, g16 = vis_net.encode(vis_input, out16_dtype=out16_dtype, precision=precision, want_f32=False)
        heads = g16.shape[1]
        g16 = g16.reshape(g16.shape[0], -1)
        return cls(g16, total, heads, rank, world_size, group, backend)

    def _all_reduce(self, t, group=None):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group if group is not None else self.group)
        return t

    # The three stages of a search.  `group` lets the pipelined path (Retriever.submit) run the collectives of the stages
    # before and after the sweep on two extra communicators, each on its own stream.
    def _gt_scores(self, q16, gt_global, group=None):
        """Raw (unscaled) score of every query's ground truth, replicated: computed by the shard that owns the ground
        truth (others contribute 0) and summed over the shards."""
        n_local = self.hi - self.lo
        owned = (gt_global >= self.lo) & (gt_global < self.hi)
        gt_local = torch.where(owned, gt_global - self.lo, torch.full_like(gt_global, -1))
        if n_local > 0:
            sgt = self.backend.gt_scores(q16, self.g16, gt_local)
        else:
            sgt = torch.zeros(q16.shape[0], dtype=torch.float32, device=q16.device)
        return self._all_reduce(sgt, group)

    def _sweep(self, q16, sgt, gt_global, k):
        """Local shard: (count of videos ranked above the ground truth, ordered top-k values, global indices)."""
        if self.hi - self.lo > 0:
            if self.timers is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            count, tv, ti = self.backend.rank_topk(q16, self.g16, sgt, gt_global, k, 1.0 / self.heads, self.lo, self._ws)
            if self.timers is not None:
                e1.record()
                self.timers.append((e0, e1))
            return count, tv, ti
        Q = q16.shape[0]
        return (torch.zeros(Q, dtype=torch.int32, device=q16.device),
                torch.full((Q, k), float("-inf"), dtype=torch.float32, device=q16.device),
                torch.full((Q, k), -1, dtype=torch.int32, device=q16.device))

    def _merge(self, count, tv, ti, k, group=None):
        """Global rank0 and top-k from the per-shard partial results: ONE collective -- [count | score bits | index] per
        query, gathered into [W, Q, 1 + 2k] -- then the counts are summed and the lists merged under the tie rule."""
        if self.world_size == 1:
            return count, tv, ti
        mine = torch.cat([count.to(torch.int32).reshape(-1, 1), tv.contiguous().view(torch.int32), ti.to(torch.int32)], 1).contiguous()
        flat = torch.empty((self.world_size * mine.shape[0], mine.shape[1]), dtype=torch.int32, device=mine.device)
        dist.all_gather_into_tensor(flat, mine, group=group if group is not None else self.group)
        both = flat.view(self.world_size, mine.shape[0], mine.shape[1])
        count = both[:, :, 0].sum(0, dtype=torch.int32)
        if k > 0:
            tv, ti = self.backend.merge(both[:, :, 1:1 + k].contiguous().view(torch.float32), both[:, :, 1 + k:].contiguous(), k)
        return count, tv, ti

    def search(self, q16: torch.Tensor, gt_global: torch.Tensor, k: int = 10) -> SearchResult:
        """q16 [Q, H*d_h] 16-bit unit-norm query embeddings (replicated on every rank), gt_global int [Q]."""
        gt_global = gt_global.to(torch.int32)
        with ops.nvtx_range("gt_scores"):
            sgt = self._gt_scores(q16, gt_global)
        with ops.nvtx_range("sweep_rank_topk"):
            part = self._sweep(q16, sgt, gt_global, k)
        with ops.nvtx_range("merge_metrics"):
            count, tv, ti = self._merge(*part, k)
            return SearchResult(count, tv, ti, self.backend.metrics(count))

    def ranked_lists(self, q16: torch.Tensor, k: int, query_chunk: int = 2048, gt_global: Optional[torch.Tensor] = None):
        """The k best videos of every query over the whole (sharded) gallery, k up to 2048: the lists the reference
        writes to id.sent.score.txt (top 2000) and t2v.pkl (top 500), predictor.py:53-88.  Queries are processed
        `query_chunk` at a time (the backend bounds its own scratch: candidate lists, or dense 256-query pieces); with W > 1 shards
        every rank extracts its local lists and one all_gather + merge per chunk yields the global ones.
        Returns (values fp32 [Q, k], global indices int32 [Q, k]) on the device, -inf / -1 beyond the gallery size.
        With gt_global (int [Q]) a third result: rank0 int32 [Q], the exact rank of every query's ground truth -- counted
        in the SAME sweep that collects the list candidates, so metrics + lists of a query set cost one pass."""
        be = self.backend
        scale = 1.0 / self.heads
        Q = q16.shape[0]
        n_local = self.hi - self.lo
        out_v, out_i, out_r = [], [], []
        want_rank = gt_global is not None
        if want_rank:
            gt_global = gt_global.to(torch.int32)
            sgt_all = self._gt_scores(q16, gt_global)
        for lo in range(0, Q, query_chunk):
            qc = q16[lo:lo + query_chunk]
            if want_rank:
                gc, sc = gt_global[lo:lo + query_chunk], sgt_all[lo:lo + query_chunk]
                if n_local > 0 and getattr(be, "lists_with_rank", False):
                    tv, ti, cnt = be.dense_topk(qc, self.g16, k, scale, self.lo, sgt=sc, gt=gc)
                    if cnt is None:
                        cnt = be.rank_topk(qc, self.g16, sc, gc, 0, scale, self.lo)[0]
                elif n_local > 0:
                    tv, ti = be.dense_topk(qc, self.g16, k, scale, self.lo)
                    cnt = be.rank_topk(qc, self.g16, sc, gc, 0, scale, self.lo)[0]
                else:
                    cnt = torch.zeros(qc.shape[0], dtype=torch.int32, device=q16.device)
                out_r.append(self._all_reduce(cnt.to(torch.int32)))
            if n_local > 0 and want_rank:
                pass
            elif n_local > 0:
                tv, ti = be.dense_topk(qc, self.g16, k, scale, self.lo)
            else:
                tv = torch.full((qc.shape[0], k), float("-inf"), dtype=torch.float32, device=q16.device)
                ti = torch.full((qc.shape[0], k), -1, dtype=torch.int32, device=q16.device)
            if self.world_size > 1:
                vals = [torch.empty_like(tv) for _ in range(self.world_size)]
                idxs = [torch.empty_like(ti) for _ in range(self.world_size)]
                dist.all_gather(vals, tv.contiguous(), group=self.group)
                dist.all_gather(idxs, ti.contiguous(), group=self.group)
                tv, ti = be.merge_lists(torch.cat(vals, 1), torch.cat(idxs, 1), k)
            out_v.append(tv)
            out_i.append(ti)
        if not out_v:
            empty = (torch.empty((0, k), dtype=torch.float32, device=q16.device), torch.empty((0, k), dtype=torch.int32, device=q16.device))
            return empty + (torch.empty(0, dtype=torch.int32, device=q16.device),) if want_rank else empty
        if want_rank:
            return torch.cat(out_v), torch.cat(out_i), torch.cat(out_r)
        return torch.cat(out_v), torch.cat(out_i)


class RankedScores:
    """What W2VVPP.predict returns in place of the dense [Q, V] score matrix when the gallery is large (the reference's
    `predict_batch` branch, model/model.py:1020-1021, :1081-1128, would build a 40 GB host matrix for 10 k x 1 M and
    re-encode the gallery for every text batch): the resident gallery index plus the fused 16-bit query embeddings,
    from which everything predictor.py:232-284 derives from the matrix is produced without materialising it --
    `search` (rank of the ground truth, top-k, R@K / MedR ... in one sweep), `ranked_lists` (the writers' top-500 /
    top-2000 lists), `rows` (a dense block of rows on the device).  `numpy()` / np.asarray() still give the dense
    matrix when it is small enough to exist."""

    dense_limit_bytes = 1 << 31

    def __init__(self, index: GalleryIndex, q16: torch.Tensor):
        self.index, self.q16 = index, q16
        self.shape = (q16.shape[0], index.total)

    def search(self, gt_global, k: int = 10) -> SearchResult:
        return self.index.search(self.q16, torch.as_tensor(gt_global).to(self.q16.device), k)

    def ranked_lists(self, k: int, query_chunk: int = 2048, gt_global=None):
        gt = None if gt_global is None else torch.as_tensor(gt_global).to(self.q16.device)
        return self.index.ranked_lists(self.q16, k, query_chunk, gt_global=gt)

    def rows(self, lo: int, hi: int) -> torch.Tensor:
        """Dense fp32 scores of queries [lo, hi) against this rank's shard, on the device."""
        return ops.sim_dense(self.q16[lo:hi], self.index.g16, 1.0 / self.index.heads)

    def numpy(self):
        if self.index.world_size > 1:
            raise ops.LaffError("RankedScores.numpy(): the gallery is sharded over %d ranks" % self.index.world_size)
        if self.shape[0] * self.shape[1] * 4 > self.dense_limit_bytes:
            raise MemoryError("the dense %d x %d score matrix is %.1f GB: use search() / ranked_lists() / rows()"
                              % (self.shape[0], self.shape[1], self.shape[0] * self.shape[1] * 4 / 1e9))
        return self.rows(0, self.shape[0]).cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)


class Retriever:
    """The whole query path behind one call: fuse the text features of a batch of queries (txt_net, F1-F6), then rank
    them against the resident gallery shard(s) (S2 + E2 + E3).  This is the public API bench.py's e2e leg times."""

    def __init__(self, txt_net, index: GalleryIndex, out16_dtype=None):
        """Queries are fused in the gallery's operand type (fp16 or bf16), projection operands included."""
        self.txt_net = txt_net
        self.index = index
        self.out16_dtype = out16_dtype or index.g16.dtype
        self.precision = "fp16" if self.out16_dtype == torch.float16 else "bf16"

    def query_slice(self, Q: int):
        """Rows [lo, hi) of a Q-query batch that this rank fuses, and the per-rank slot size of the all-gather."""
        W, r = self.index.world_size, self.index.rank
        per = (Q + W - 1) // W
        return min(Q, r * per), min(Q, (r + 1) * per), per

    @torch.no_grad()
    def encode_queries(self, caption_feat_dict, total: Optional[int] = None, group=None) -> torch.Tensor:
        """Fused 16-bit query embeddings [Q, H*d_h], replicated on every rank.  With W > 1 ranks each rank fuses only
        its 1/W slice of the queries (the fusion is data-parallel per item) and the slices are all-gathered over
        NVLink, so the replicated part of a step shrinks with W.  `total` = Q says that caption_feat_dict already holds
        only this rank's slice (query_slice(Q)) of a Q-query batch -- host callers then copy 1/W of the features."""
        idx = self.index
        W = idx.world_size
        if W == 1:
            with ops.nvtx_range("fuse_queries"):
                _, q16 = self.txt_net.encode(caption_feat_dict, out16_dtype=self.out16_dtype, precision=self.precision)
            return q16.reshape(q16.shape[0], -1)
        if total is None:
            Q = next(iter(caption_feat_dict.values())).shape[0]
            lo, hi, per = self.query_slice(Q)
            part = {k: v[lo:hi] for k, v in caption_feat_dict.items()}
        else:
            Q = int(total)
            lo, hi, per = self.query_slice(Q)
            part = caption_feat_dict
            if next(iter(part.values())).shape[0] != hi - lo:
                raise ValueError("encode_queries(total=%d): expected this rank's %d rows" % (Q, hi - lo))
        D = idx.g16.shape[1]
        buf = torch.zeros((per, D), dtype=self.out16_dtype, device=idx.g16.device)
        if hi > lo:
            with ops.nvtx_range("fuse_queries"):
                _, q16 = self.txt_net.encode(part, out16_dtype=self.out16_dtype, precision=self.precision)
            buf[: hi - lo] = q16.reshape(hi - lo, -1)
        out = torch.empty((W * per, D), dtype=buf.dtype, device=buf.device)
        with ops.nvtx_range("all_gather_queries"):
            dist.all_gather_into_tensor(out, buf, group=group if group is not None else idx.group)
        return out[:Q]

    @torch.no_grad()
    def ranked_lists(self, caption_feat_dict, k: int, query_chunk: int = 256):
        """Fuse the queries, then GalleryIndex.ranked_lists (the writer lists of predictor.py:53-88 at gallery scale)."""
        return self.index.ranked_lists(self.encode_queries(caption_feat_dict), k, query_chunk)

    @torch.no_grad()
    def rank(self, caption_feat_dict, gt_global, k: int = 10, chunks: int = 1) -> SearchResult:
        """caption_feat_dict: per-encoder text features (host or device tensors); gt_global: int [Q].

        chunks > 1 with host (pinned) inputs: the queries are cut into `chunks` pieces whose host->device copies are
        issued up front on a copy stream; piece i is fused and swept while pieces i+1.. are still in flight, so only
        the first piece's copy is exposed.  Results are identical to chunks = 1 (queries are independent)."""
        first = next(iter(caption_feat_dict.values()))
        Q = first.shape[0]
        if chunks <= 1 or first.is_cuda or Q < 2 * chunks:
            q16 = self.encode_queries(caption_feat_dict)
            return self.index.search(q16, gt_global.to(q16.device, non_blocking=True), k)
        dev = self.index.g16.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        cs.wait_stream(main)
        per = (Q + chunks - 1) // chunks
        staged = []
        with torch.cuda.stream(cs):
            for lo in range(0, Q, per):
                hi = min(Q, lo + per)
                a, b, _ = self.query_slice(hi - lo)         # this rank fuses rows [lo + a, lo + b) of the piece
                part = {name: v[lo + a:lo + b].to(dev, non_blocking=True) for name, v in caption_feat_dict.items()}
                g = gt_global[lo:hi].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                staged.append((part, g, ev, hi - lo))
        outs = []
        for part, g, ev, n in staged:
            main.wait_event(ev)
            for t in list(part.values()) + [g]:
                t.record_stream(main)
            q16 = self.encode_queries(part, total=n)
            res = self.index.search(q16, g, k)
            outs.append(res)
        rank0 = torch.cat([r.rank0 for r in outs])
        return SearchResult(rank0, torch.cat([r.topk_val for r in outs]), torch.cat([r.topk_idx for r in outs]),
                            self.index.backend.metrics(rank0))

    # ------------------------------------------------------------------------------------------------------------
    # latency path: small batches are launch-bound (C2, 2990 x 2990: ~40 launches for ~0.2 ms of tensor work)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def rank_graphed(self, caption_feat_dict, gt_global, k: int = 10) -> SearchResult:
        """rank() replayed as ONE CUDA graph: the first call with a given set of input shapes runs eagerly once (lazy
        operand caches), captures fuse -> ground-truth scores -> sweep -> metrics, and every later call with those shapes
        copies its inputs into the graph's static buffers and replays it.  One process (no collectives), device-resident
        inputs.  The returned tensors are the graph's static outputs: read them before the next call with these shapes."""
        idx = self.index
        if idx.world_size != 1:
            raise ops.LaffError("rank_graphed: one process per gallery (use submit() with several ranks)")
        dev = idx.g16.device
        feats = {n: v.to(dev) for n, v in caption_feat_dict.items()}
        gt = gt_global.to(dev)

        def sig_of(v):
            if isinstance(v, ops.SparseRows):
                return ("csr", v.shape, v.ids.numel(), v.base)
            return (tuple(v.shape), str(v.dtype))
        sig = tuple((n, sig_of(v)) for n, v in feats.items()) + (k, tuple(gt.shape), str(gt.dtype))
        cache = self.__dict__.setdefault("_rank_graphs", {})
        g = cache.get(sig)
        if g is None:
            def clone(v):
                if isinstance(v, ops.SparseRows):
                    return ops.SparseRows(v.offsets.clone(), v.ids.clone(), v.ndims, v.base,
                                          None if v.row_scale is None else v.row_scale.clone())
                return v.clone()
            static = {n: clone(v) for n, v in feats.items()}
            sgt = gt.clone()
            self.rank(static, sgt, k)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                res = self.rank(static, sgt, k)
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            g = cache[sig] = (graph, static, sgt, res)
        graph, static, sgt, res = g
        for n, v in feats.items():
            if isinstance(v, ops.SparseRows):
                static[n].offsets.copy_(v.offsets, non_blocking=True)
                static[n].ids.copy_(v.ids, non_blocking=True)
            else:
                static[n].copy_(v, non_blocking=True)
        sgt.copy_(gt, non_blocking=True)
        graph.replay()
        return res

    # ------------------------------------------------------------------------------------------------------------
    # pipelined path
    # ------------------------------------------------------------------------------------------------------------
    # Measured at 4 and 8 GPUs (profiles/r02_scale_experiments.md).  The side stages need SMs of their own while a sweep runs:
    # a post-stage all-gather spins until the slowest rank's sweep is done, and if it (plus the pre-stage collective) leaves
    # no SM for the pre stage's compute kernels, every rank's next sweep waits (reserve 2 / 1 CTA: 12.3 ms per step at 8
    # GPUs instead of 9.1 serial).  4 reserved SMs with 2-CTA collectives is the measured optimum (8.31 ms); 6 and 8 cost
    # the sweep more than they save (72 CTA pairs still cover a launch of 4890 units in 68 waves, 71 need 69).
    reserve_sms = 4      # SMs left to the side streams while a sweep runs with W > 1 ranks (W = 1: none, see _pipe)
    side_max_ctas = 2    # CTAs of an NCCL kernel on the side communicators

    _side_groups: dict = {}   # (ranks, max_ctas, stage) -> communicator, shared by every Retriever of the process

    def _side_group(self, stage: str):
        """A communicator over the index's ranks for one side stream; NCCL kernels limited to `side_max_ctas` CTAs so that
        the collective of the stage before a sweep and the one after it fit the reserved SMs together (a kernel waiting
        for its peers then never blocks the other stage's).  Created once per (ranks, CTA limit, stage) and reused:
        new_group is a collective call and communicators are not free."""
        idx = self.index
        ranks = tuple(dist.get_process_group_ranks(idx.group) if idx.group is not None else range(dist.get_world_size()))
        key = (ranks, int(self.side_max_ctas), stage)
        g = Retriever._side_groups.get(key)
        if g is None:
            if dist.get_backend(idx.group) == "nccl":
                opts = dist.ProcessGroupNCCL.Options()
                try:
                    opts.config.max_ctas = int(self.side_max_ctas)
                    opts.config.min_ctas = 1
                except Exception:
                    pass
                g = dist.new_group(list(ranks), pg_options=opts)
            else:
                g = dist.new_group(list(ranks))
            Retriever._side_groups[key] = g
        return g

    def _pipe(self):
        P = getattr(self, "_pipe_state", None)
        if P is not None:
            return P
        idx = self.index
        dev = idx.g16.device
        P = types.SimpleNamespace(pre=None, sweep=None, post=None, copy=None, g_pre=idx.group, g_post=idx.group,
                                  sweep_sms=0, side_sms=0)
        if idx.world_size > 1:                                   # collective calls: every rank creates them in this order
            P.g_pre, P.g_post = self._side_group("pre"), self._side_group("post")
        if dev.type == "cuda":
            P.pre, P.post, P.copy = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            P.sweep = torch.cuda.Stream(dev, priority=-1)        # a sweep launch gets freed SMs before queued side work
            total = torch.cuda.get_device_properties(dev).multi_processor_count
            # one rank: no collectives, the side stages are short kernels that fill the gaps between sweep launches
            r = 0 if idx.world_size == 1 and not getattr(self, "reserve_on_one_gpu", False) else max(0, min(int(self.reserve_sms), total - 2))
            r += r % 2                                           # CTA pairs
            P.side_sms, P.sweep_sms = (r, total - r) if r > 0 else (0, 0)
        self._pipe_state = P
        return P

    @torch.no_grad()
    def submit(self, caption_feat_dict, gt_global, k: int = 10, pieces: Optional[int] = None, inputs_ready=None,
               fetch: bool = False) -> PendingSearch:
        """rank() as a three-stage pipeline over pieces of the query batch, each stage on its own stream:

            pre   fuse this rank's slice of the piece, all-gather the embeddings, ground-truth scores (+ all-reduce)
            sweep similarity + rank + top-k over the local gallery shard (high-priority stream, all but `reserve_sms` SMs)
            post  one all-gather of [count | lists], merge; after the last piece: metrics (and the packed D2H if fetch)

        The stages of consecutive pieces -- and of consecutive submits, which need not wait for each other -- overlap:
        the pre / post kernels and the NCCL kernels of their collectives (separate communicators, `side_max_ctas` CTAs)
        run on the reserved SMs while the sweep of another piece owns the rest, so neither their run time nor the
        rank-to-rank skew of a collective stalls the tensor cores.  Results are identical to rank(): queries are
        independent and every stage is the same kernel sequence.

        inputs_ready: None -- the inputs are complete once the work already queued on the caller's current stream has run
        (an event is recorded there); a torch.cuda.Event to wait for; or False: they are complete now (resident inputs).
        Host (pinned) inputs are copied piece by piece on a copy stream."""
        idx = self.index
        P = self._pipe()
        dev = idx.g16.device
        first = next(iter(caption_feat_dict.values()))
        Q = first.shape[0]
        cuda = dev.type == "cuda"
        if Q == 0:
            raise ValueError("submit: empty query batch")
        if pieces is None:                                       # the sweep's own row groups (10 row tiles of 256), at most 4;
            pieces = max(1, min(4 if idx.world_size == 1 else 2, (Q + 2559) // 2560))   # with collectives per piece: 2
        pieces = max(1, min(int(pieces), max(1, Q)))
        per = (Q + pieces - 1) // pieces
        per = (per + 255) // 256 * 256 if per > 256 else per     # whole row tiles per piece
        host_in = cuda and not first.is_cuda
        stream = (lambda s: torch.cuda.stream(s)) if cuda else (lambda s: contextlib.nullcontext())
        dep = None
        if cuda and not host_in and inputs_ready is not False:
            dep = inputs_ready
            if dep is None:
                dep = torch.cuda.Event()
                dep.record(torch.cuda.current_stream(dev))
        # ---- stage the pieces (host inputs: H2D copies of this rank's slice of every piece, issued up front)
        staged = []
        for lo in range(0, Q, per):
            hi = min(Q, lo + per)
            a, b, _ = self.query_slice(hi - lo)                  # this rank fuses rows [lo + a, lo + b) of the piece
            if host_in:
                with stream(P.copy):
                    part = {n: v[lo + a:lo + b].to(dev, non_blocking=True) for n, v in caption_feat_dict.items()}
                    g = gt_global[lo:hi].to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(P.copy)
            else:
                part = {n: v[lo + a:lo + b] for n, v in caption_feat_dict.items()}
                g = gt_global[lo:hi].to(dev) if not gt_global.is_cuda and cuda else gt_global[lo:hi]
                ev = dep
            staged.append((part, g, ev, hi - lo))
        outs = []
        for part, g, ev, n in staged:
            with stream(P.pre), ops.sm_limit(P.side_sms) if cuda else contextlib.nullcontext():
                if cuda and ev is not None:
                    P.pre.wait_event(ev)
                if host_in:
                    for t in list(part.values()) + [g]:
                        t.record_stream(P.pre)
                q16 = self.encode_queries(part, total=n, group=P.g_pre) if idx.world_size > 1 else self.encode_queries(part)
                g32 = g.to(torch.int32)
                with ops.nvtx_range("gt_scores"):
                    sgt = idx._gt_scores(q16, g32, group=P.g_pre)
                if cuda:
                    ready = torch.cuda.Event()
                    ready.record(P.pre)
            with stream(P.sweep), ops.sm_limit(P.sweep_sms) if cuda else contextlib.nullcontext():
                if cuda:
                    P.sweep.wait_event(ready)
                    for t in (q16, g32, sgt):
                        t.record_stream(P.sweep)
                with ops.nvtx_range("sweep_rank_topk"):
                    count, tv, ti = idx._sweep(q16, sgt, g32, k)
                if cuda:
                    swept = torch.cuda.Event()
                    swept.record(P.sweep)
            with stream(P.post), ops.sm_limit(P.side_sms) if cuda else contextlib.nullcontext():
                if cuda:
                    P.post.wait_event(swept)
                    for t in (count, tv, ti):
                        t.record_stream(P.post)
                with ops.nvtx_range("merge"):
                    outs.append(idx._merge(count, tv, ti, k, group=P.g_post))
        with stream(P.post), ops.sm_limit(P.side_sms) if cuda else contextlib.nullcontext():
            rank0 = outs[0][0] if len(outs) == 1 else torch.cat([o[0] for o in outs])
            tv = outs[0][1] if len(outs) == 1 else torch.cat([o[1] for o in outs])
            ti = outs[0][2] if len(outs) == 1 else torch.cat([o[2] for o in outs])
            res = SearchResult(rank0, tv, ti, idx.backend.metrics(rank0))
            packed = res._pack_to_pinned() if (fetch and cuda) else None
            done = None
            if cuda:
                done = torch.cuda.Event()
                done.record(P.post)
        return PendingSearch(res, done, P.post, packed)
