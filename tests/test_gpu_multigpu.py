"""Multi-GPU parity as a collected test: tools/check_multigpu.py under torchrun on 2 GPUs of the box -- NCCL,
the real kernels, gallery sharded over the ranks -- must reproduce the single-process answer bit for bit: the serial
search, the writer lists, the sharded query fusion with chunked host copies, and the pipelined submit path with its two
extra communicators.  Skipped on a box with fewer than 2 GPUs (the CPU stand-in is tests/test_retrieval_gloo_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box")
def test_two_gpu_search_is_bit_identical_to_one_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "check_multigpu.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU parity world=2: OK" in r.stdout and "pipelined submit" in r.stdout and "MISMATCH" not in r.stdout
