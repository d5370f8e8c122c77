"""Multi-GPU parity of the gallery-sharded search (run under torchrun on 2/4/8 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tools/check_multigpu.py

Every rank builds the same synthetic problem, holds its shard, and runs GalleryIndex.search over NCCL; rank 0 also runs
the single-shard search on the whole gallery and requires bit-identical rank0 / top-k / metrics.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from laff_b200 import synth  # noqa: E402
from laff_b200.retrieval import GalleryIndex, Retriever, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    Q, V, H, dh, k = 1500, 100003, 8, 512, 10
    gen = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(V, H, dh, generator=gen, device=dev)
    g16 = (x / x.norm(dim=2, keepdim=True)).reshape(V, -1).to(torch.bfloat16)
    gt = (torch.arange(Q, device=dev) * 97) % V
    n = torch.randn(Q, H, dh, generator=gen, device=dev)
    qf = g16[gt].float().view(Q, H, dh) + synth.sigma_for_recall(V, H * dh) * n / n.norm(dim=2, keepdim=True)
    q16 = (qf / qf.norm(dim=2, keepdim=True)).reshape(Q, -1).to(torch.bfloat16)
    g16[V - 1] = g16[gt[3]]
    g16[7] = g16[gt[9]]
    lo, hi = shard_bounds(V, world, rank)
    res = GalleryIndex(g16[lo:hi].contiguous(), V, H, rank, world).search(q16, gt.to(torch.int32), k)
    lv, li = GalleryIndex(g16[lo:hi].contiguous(), V, H, rank, world).ranked_lists(q16[:300], 500, query_chunk=128)
    # lists + exact ranks from one sweep per shard (laff_sim_collect_rank), counts summed over the shards
    lv_r, li_r, rk_r = GalleryIndex(g16[lo:hi].contiguous(), V, H, rank, world).ranked_lists(q16[:300], 500, query_chunk=128,
                                                                                         gt_global=gt[:300].to(torch.int32))
    # the public query path: pinned host features in 3 pieces, each rank copies and fuses only its slice of every piece
    import bench  # noqa: E402  (synthetic text net + query features of the bench)
    txt_net = bench.build_txt_net(dev)
    feats = bench.query_features(1001, pinned=True)
    gt2 = ((torch.arange(1001) * 97) % V).to(torch.int32)
    rr = Retriever(txt_net, GalleryIndex(g16[lo:hi].contiguous(), V, H, rank, world)).rank(feats, gt2.pin_memory(), k, chunks=3)
    # the pipelined path: pieces through pre / sweep / post stages on their own streams and communicators, two submits in
    # flight (device-resident inputs, then pinned host inputs with the packed D2H enqueued at submit time)
    retr_p = Retriever(txt_net, GalleryIndex(g16[lo:hi].contiguous(), V, H, rank, world))
    feats_dev = {n: v.to(dev) for n, v in feats.items()}
    p1 = retr_p.submit(feats_dev, gt2.to(dev), k, pieces=3, inputs_ready=None)
    p2 = retr_p.submit(feats, gt2.pin_memory(), k, pieces=2, fetch=True)
    pr1, pr2 = p1.result(), p2.to_host()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        r1 = Retriever(txt_net, GalleryIndex(g16, V, H)).rank({n: v.to(dev) for n, v in feats.items()}, gt2.to(dev), k)
        ok = torch.equal(rr.rank0, r1.rank0) and torch.equal(rr.topk_idx, r1.topk_idx) and torch.equal(rr.topk_val, r1.topk_val)
        print("sharded query fusion + chunked host copies vs single process: %s" % ("OK" if ok else "MISMATCH"), flush=True)
        okp = (torch.equal(pr1.rank0, r1.rank0) and torch.equal(pr1.topk_idx, r1.topk_idx) and torch.equal(pr1.topk_val, r1.topk_val)
               and torch.equal(pr1.metrics, r1.metrics) and torch.equal(pr2.rank0, r1.rank0.cpu()) and torch.equal(pr2.topk_idx, r1.topk_idx.cpu())
               and torch.equal(pr2.topk_val, r1.topk_val.cpu()) and torch.equal(pr2.metrics, r1.metrics.cpu()))
        print("pipelined submit (3 + 2 pieces, two in flight) vs single process: %s" % ("OK" if okp else "MISMATCH"), flush=True)
        ok = ok and okp
    if rank == 0:
        single = GalleryIndex(g16, V, H)
        ref = single.search(q16, gt.to(torch.int32), k)
        rv, ri = single.ranked_lists(q16[:300], 500, query_chunk=128)
        ok = ok and (torch.equal(res.rank0, ref.rank0) and torch.equal(res.topk_idx, ref.topk_idx)
              and torch.equal(res.topk_val, ref.topk_val) and torch.equal(res.metrics, ref.metrics)
              and torch.equal(li, ri) and torch.equal(lv, rv) and torch.equal(ri[:, :k], ref.topk_idx[:300])
              and torch.equal(li_r, ri) and torch.equal(lv_r, rv) and torch.equal(rk_r, ref.rank0[:300]))
        print("multi-GPU parity world=%d: %s  R@1=%.2f R@10=%.2f MedR=%.0f" % (
            world, "OK" if ok else "MISMATCH", ref.metrics[0].item(), ref.metrics[2].item(), ref.metrics[3].item()), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
