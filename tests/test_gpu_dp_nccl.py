"""Data-parallel step on real GPUs (NCCL): runs tools/check_dp.py under torchrun when the box has
at least two GPUs (the single-GPU round-end tier skips it; tests/test_dp_gloo.py covers the host
logic on CPU).  The tool compares the all-reduced gradients of the sharded step -- eager
event-driven overlap, the same step replayed as one CUDA graph with the NCCL all-reduces captured,
phase-wise overlap and no overlap -- with the whole global batch on one GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_dp_two_gpus_matches_single_gpu(mode):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541" if mode == "bf16" else "29542", os.path.join(ROOT, "tools", "check_dp.py"), "--mode", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("dp world=2") == 4, out.stdout


@pytest.mark.parametrize("comm", ["fp32", "bf16"])
def test_dp_two_gpus_late_schedule(comm):
    """GradReducer(schedule="late"): one all-reduce of the flat gradient buffer after the backward; eager and graph."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29544" if comm == "fp32" else "29545", os.path.join(ROOT, "tools", "check_dp.py"), "--mode", "bf16",
           "--grad-comm", comm, "--reduce", "late"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("dp world=2") == 2, out.stdout


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_dp_two_gpus_multimem_allreduce(mode):
    """GradReducer(schedule="multimem"): gradients in symmetric memory, one mmqg_allreduce_multimem launch per group (barrier
    over the peers' signal pads, multimem.ld_reduce / multimem.st over NVSwitch multicast memory); eager and graph."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29546" if mode == "fp32" else "29547", os.path.join(ROOT, "tools", "check_dp.py"), "--mode", mode,
           "--reduce", "multimem"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("dp world=2") == 2, out.stdout


def test_dp_two_gpus_bf16_gradient_exchange():
    """GradReducer(comm_dtype=bfloat16): buckets packed to bf16 (mmqg_pack_bf16), all-reduced, unpacked; eager and graph."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tools", "check_dp.py"), "--mode", "bf16", "--grad-comm", "bf16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("dp world=2") == 2, out.stdout
