"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/mmqg.h declares; host-side argument validation works without a GPU;
the product package never imports the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    from mmqg import build, _cabi
    build.build()
    return _cabi


def test_header_symbols_all_exported(cabi):
    hdr = open(os.path.join(ROOT, "include", "mmqg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mmqg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = cabi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in mmqg.h but not exported"
    assert declared == set(cabi.SYMBOLS), declared ^ set(cabi.SYMBOLS)
    assert L.mmqg_abi_version() == 3


def test_struct_layouts(cabi):
    assert C.sizeof(cabi.MmqgDims) == 13 * 4
    assert C.sizeof(cabi.MmqgTensors) == 8 * (1 + 4 * 4 + 4 + 6 + 4 * 4 + 2)
    assert C.sizeof(cabi.MmqgBatch) == 56      # 4 tensors + 3 optional length arrays


def test_workspace_and_validation_without_gpu(cabi):
    from mmqg.dims import config
    L = cabi.lib()
    d = cabi.c_dims(config(2))
    n = L.mmqg_train_workspace_bytes(C.byref(d), 0)
    assert 1 << 30 < n < 8 << 30          # ~2 GB of fp32 activations at cfg-2
    bad = cabi.c_dims(config(2))
    bad.T_t = 500                          # > TM
    assert L.mmqg_train_workspace_bytes(C.byref(bad), 0) == 0
    assert b"T_t<=TM" in L.mmqg_last_error()
    bad.T_t, bad.L = 100, 9
    assert L.mmqg_train_workspace_bytes(C.byref(bad), 0) == 0
    # null tensors are rejected before any CUDA call
    t = cabi.MmqgTensors()
    b = cabi.MmqgBatch()
    st = L.mmqg_train_forward(C.byref(d), C.byref(t), C.byref(b), None, 0, None, 0, None, 1.0, 0.0, 0, 0, None)
    assert st == 1 and b"null" in L.mmqg_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multi-modal-qg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b|oracle/|import_module\(.oracle", src, re.M), \
                    os.path.join(dirpath, f)


def test_engine_refuses_cpu(cabi):
    import torch
    from mmqg.dims import Dims
    from mmqg.synth import make_params
    from mmqg.engine import TrainEngine
    d = Dims(B=1, T_t=2, T_v=1, T_q=2, V=11, E=4, H=8, L=1, H_a=4, H_v=8, F_v=4, TM=3, AM=2)
    with pytest.raises(cabi.MmqgError):
        TrainEngine(d, make_params(d), device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            TrainEngine(d, make_params(d), device="cuda")
