"""Error behaviour of the C ABI on a GPU box: bad arguments come back as LAFF_E* codes with a message, never as a
crash or a silent fallback."""
import ctypes as C

import pytest
import torch

from laff_b200 import LaffError, _capi, ops

pytestmark = pytest.mark.gpu


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_operand_validation():
    q = torch.zeros(8, 64, dtype=torch.bfloat16, device="cuda")
    g = torch.zeros(16, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(LaffError, match="fp16 or bf16"):
        ops.sim_dense(q.float(), g.float())
    with pytest.raises(LaffError, match="K mismatch"):
        ops.sim_dense(q, g[:, :32].contiguous())
    with pytest.raises(LaffError, match="CUDA tensors"):
        ops.sim_dense(q.cpu(), g.cpu())
    lib = _capi.lib()
    out = torch.zeros(8, 16, device="cuda")
    # K not a multiple of 8 -> LAFF_EINVAL with a message
    rc = lib.laff_sim_dense(q.data_ptr(), g.data_ptr(), 8, 16, 60, 64, 64, 1, 1.0, out.data_ptr(), 16, _st())
    assert rc == -1 and b"multiples of 8" in lib.laff_last_error()
    # ld_out < V
    rc = lib.laff_sim_dense(q.data_ptr(), g.data_ptr(), 8, 16, 64, 64, 64, 1, 1.0, out.data_ptr(), 8, _st())
    assert rc == -1
    # null pointer
    rc = lib.laff_sim_dense(None, g.data_ptr(), 8, 16, 64, 64, 64, 1, 1.0, out.data_ptr(), 16, _st())
    assert rc == -1 and b"null" in lib.laff_last_error()


def test_rank_topk_limits_and_workspace():
    q = torch.zeros(8, 64, dtype=torch.bfloat16, device="cuda")
    g = torch.zeros(300, 64, dtype=torch.bfloat16, device="cuda")
    gt = torch.zeros(8, dtype=torch.int32, device="cuda")
    sgt = torch.zeros(8, device="cuda")
    with pytest.raises(LaffError, match=r"k=17 outside"):
        ops.sim_rank_topk(q, g, sgt, gt, 17)
    lib = _capi.lib()
    cnt = torch.zeros(8, dtype=torch.int32, device="cuda")
    tv = torch.zeros(8, 4, device="cuda")
    ti = torch.zeros(8, 4, dtype=torch.int32, device="cuda")
    ws = torch.zeros(64, dtype=torch.uint8, device="cuda")
    rc = lib.laff_sim_rank_topk(q.data_ptr(), g.data_ptr(), 8, 300, 64, 64, 64, 1, 1.0, sgt.data_ptr(), gt.data_ptr(), 0, 4,
                                cnt.data_ptr(), tv.data_ptr(), ti.data_ptr(), ws.data_ptr(), 64, _st())
    assert rc == -4 and b"workspace too small" in lib.laff_last_error()
    # all-zero operands: every score ties with s_gt = 0; the tie rule still gives a definite answer
    c, v, i = ops.sim_rank_topk(q, g, sgt, gt, 4)
    assert c.tolist() == [299] * 8                      # every other video has a higher index than gt = 0
    assert i[0].tolist() == [299, 298, 297, 296] and v[0].tolist() == [0.0] * 4


def test_pool_and_loss_validation():
    w = torch.zeros(8, 32, device="cuda")
    b = torch.zeros(8, device="cuda")
    y = torch.zeros(4, 256, device="cuda")
    with pytest.raises(LaffError, match="n_features"):
        ops.attention_pool([{"y": y}] * 9, w, b, 8, 32)
    with pytest.raises(LaffError, match="multiple of 32"):
        ops.attention_pool([{"y": torch.zeros(4, 8 * 20, device="cuda")}], torch.zeros(8, 20, device="cuda"), b, 8, 20)
    with pytest.raises(LaffError, match="must divide D"):
        ops.attention_pool([{"x": torch.zeros(4, 48, device="cuda")}], w, b, 8, 32)
    with pytest.raises(KeyError):
        ops.mrl_forward_backward(torch.zeros(4, 2, 8, device="cuda"), torch.zeros(4, 2, 8, device="cuda"), 0.2, True, "sideways", "sum")
    # empty batches are a no-op, not an error (the reference's loaders can yield a short last batch)
    assert ops.l2norm_quantize(torch.zeros(0, 64, device="cuda"), 2).shape == (0, 64)
    assert ops.sim_dense(torch.zeros(0, 64, dtype=torch.bfloat16, device="cuda"),
                         torch.zeros(5, 64, dtype=torch.bfloat16, device="cuda")).shape == (0, 5)
