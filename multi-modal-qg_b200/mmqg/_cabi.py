"""ctypes binding of libmmqg.so (include/mmqg.h).  No CPU fallback: if the library is
missing or the device is not sm_100 every call raises."""
import ctypes as C
import os

from .dims import Dims

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMQG_LIB") or os.path.join(HERE, "libmmqg.so")      # MMQG_LIB: an alternative in-tree build (A/B runs)
MAX_LAYERS = 4
MODE_FP32, MODE_BF16, MODE_FP32_TC = 0, 1, 2

_fp = C.c_void_p      # device pointers travel as integers (tensor.data_ptr())


class MmqgDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "T_t", "T_v", "T_q", "V", "E", "H", "L", "H_a", "H_v", "F_v", "TM", "AM")]


class MmqgTensors(C.Structure):
    _fields_ = [("emb", _fp),
                ("text_w_ih", _fp * MAX_LAYERS), ("text_w_hh", _fp * MAX_LAYERS),
                ("text_b_ih", _fp * MAX_LAYERS), ("text_b_hh", _fp * MAX_LAYERS),
                ("vid_w_ih", _fp), ("vid_w_hh", _fp), ("vid_b_ih", _fp), ("vid_b_hh", _fp),
                ("attn_w", _fp * 3), ("attn_b", _fp * 3),
                ("dec_w_ih", _fp * MAX_LAYERS), ("dec_w_hh", _fp * MAX_LAYERS),
                ("dec_b_ih", _fp * MAX_LAYERS), ("dec_b_hh", _fp * MAX_LAYERS),
                ("out_w", _fp), ("out_b", _fp)]


class MmqgBatch(C.Structure):
    _fields_ = [("context", _fp), ("target", _fp), ("frames", _fp), ("audio", _fp),
                ("ctx_len", _fp), ("tgt_len", _fp), ("n_frames", _fp)]


class MmqgGemmArgs(C.Structure):
    _fields_ = [("A", _fp), ("B", _fp), ("lda", C.c_int), ("ldb", C.c_int), ("K", C.c_int),
                ("A2", _fp), ("B2", _fp), ("lda2", C.c_int), ("ldb2", C.c_int), ("K2", C.c_int),
                ("C", _fp), ("ldc", C.c_int), ("Cin", _fp), ("ldcin", C.c_int), ("bias", _fp),
                ("M", C.c_int), ("N", C.c_int), ("alpha", C.c_float), ("beta", C.c_float),
                ("transA", C.c_int), ("transB", C.c_int), ("split_k", C.c_int), ("c_split_stride", C.c_longlong)]


class MmqgGemmBf16Args(C.Structure):
    _fields_ = [("A", _fp), ("B", _fp), ("lda", C.c_int), ("ldb", C.c_int), ("K", C.c_int),
                ("A2", _fp), ("B2", _fp), ("lda2", C.c_int), ("ldb2", C.c_int), ("K2", C.c_int),
                ("a_mn_major", C.c_int), ("b_mn_major", C.c_int),
                ("C", _fp), ("ldc", C.c_int), ("c_bf16", C.c_int), ("Cin", _fp), ("ldcin", C.c_int), ("bias", _fp),
                ("M", C.c_int), ("N", C.c_int), ("alpha", C.c_float), ("beta", C.c_float),
                ("split_k", C.c_int), ("c_split_stride", C.c_longlong)]


# every symbol include/mmqg.h declares: name -> (restype, argtypes)
_i, _ll, _f, _sz, _ull = C.c_int, C.c_longlong, C.c_float, C.c_size_t, C.c_ulonglong
_P = C.POINTER
SYMBOLS = {
    "mmqg_abi_version": (_i, []),
    "mmqg_last_error": (C.c_char_p, []),
    "mmqg_device_ok": (_i, [_i]),
    "mmqg_launch_count": (_ull, []),
    "mmqg_probe_start": (_i, [_i]),
    "mmqg_probe_stop": (_i, [_P(C.c_double), _P(_ull), _P(C.c_double), _P(C.c_double)]),
    "mmqg_train_workspace_bytes": (_sz, [_P(MmqgDims), _i]),
    "mmqg_train_forward": (_i, [_P(MmqgDims), _P(MmqgTensors), _P(MmqgBatch), _fp, _sz, _fp, _i, _P(MmqgTensors),
                                _f, _f, _ull, _i, _fp]),
    "mmqg_train_backward": (_i, [_P(MmqgDims), _P(MmqgTensors), _P(MmqgBatch), _fp, _sz, _P(MmqgTensors), _i, _f,
                                 _ull, _i, _fp]),
    "mmqg_train_backward_events": (_i, [_P(MmqgDims), _P(MmqgTensors), _P(MmqgBatch), _fp, _sz, _P(MmqgTensors),
                                        _P(_fp), _f, _ull, _i, _fp]),
    "mmqg_greedy_workspace_bytes": (_sz, [_P(MmqgDims), _i, _i]),
    "mmqg_greedy_decode": (_i, [_P(MmqgDims), _P(MmqgTensors), _P(MmqgBatch), _fp, _sz, _fp, _i, _i, _fp]),
    "mmqg_sample_decode": (_i, [_P(MmqgDims), _P(MmqgTensors), _P(MmqgBatch), _fp, _sz, _fp, _i, _ull, _i, _fp]),
    "mmqg_sample_rows": (_i, [_fp, _i, _fp, _ll, _i, _i, _ull, _ull, _fp]),
    "mmqg_sample_uniform": (_i, [_fp, _i, _ull, _ull, _fp]),
    "mmqg_dropout_mask": (_i, [_fp, _ll, _ull, _i, _f, _fp]),
    "mmqg_adam_step": (_i, [_fp, _fp, _fp, _fp, _ll, _ll, _ll, _f, _f, _f, _f, _fp, _fp]),
    "mmqg_gemm_f32": (_i, [_P(MmqgGemmArgs), _fp]),
    "mmqg_gemm_bf16": (_i, [_P(MmqgGemmBf16Args), _fp]),
    "mmqg_embedding_gather": (_i, [_fp, _fp, _fp, _i, _i, _i, _fp]),
    "mmqg_embedding_scatter_add": (_i, [_fp, _fp, _fp, _i, _i, _i, _fp]),
    "mmqg_lstm_pointwise_fwd": (_i, [_fp, _i, _fp, _i, _fp, _i, _fp, _i, _fp, _i, _i, _i, _fp]),
    "mmqg_lstm_pointwise_bwd": (_i, [_fp, _i, _fp, _i, _fp, _i, _fp, _i, _i, _ll, _fp, _i, _i, _ll, _fp, _i,
                                     _fp, _i, _i, _i, _i, _fp]),
    "mmqg_attn_fwd": (_i, [_fp, _i, _fp, _fp, _fp, _fp, _i] + [_i] * 8 + [_fp]),
    "mmqg_attn_bwd": (_i, [_fp, _i, _fp, _i, _fp, _fp, _fp, _fp, _fp] + [_i] * 8 + [_fp]),
    "mmqg_nll_rows": (_i, [_fp, _i, _fp, _ll, _fp, _i, _i, _f, _fp]),
    "mmqg_argmax_rows": (_i, [_fp, _i, _fp, _ll, _i, _i, _fp]),
    "mmqg_colsum": (_i, [_fp, _i, _fp, _i, _i, _f, _fp]),
    "mmqg_pack_bf16": (_i, [_fp, _fp, _ll, _fp]),
    "mmqg_conv_relu_fwd": (_i, [_fp] * 7 + [_i] * 7 + [_fp]),
    "mmqg_conv_stats_parts": (_i, [_i] * 5),
    "mmqg_bn_finalize": (_i, [_fp, _i, _ll, _fp, _fp, _f, _f] + [_fp] * 6 + [_i, _fp]),
    "mmqg_bn_maxpool_fwd": (_i, [_fp] * 5 + [_i] * 5 + [_fp]),
    "mmqg_maxpool_bwd": (_i, [_fp] * 3 + [_i] * 5 + [_fp]),
    "mmqg_bn_relu_bwd": (_i, [_fp] * 7 + [_i] * 4 + [_fp]),
    "mmqg_allreduce_multimem": (_i, [_fp, _ll, _fp, _i, _i, _i, _fp]),
    "mmqg_bn_relu_pool_bwd": (_i, [_fp] * 6 + [_i] + [_fp] * 2 + [_i] * 4 + [_fp]),
    "mmqg_conv_bwd_w": (_i, [_fp] * 6 + [_i] * 7 + [_fp]),
    "mmqg_conv_bwd_x": (_i, [_fp] * 3 + [_i] * 7 + [_fp]),
    "mmqg_unpack_bf16": (_i, [_fp, _fp, _ll, _fp]),
    "mmqg_lstm_seq_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mmqg_lstm_seq_fwd": (_i, [_fp] * 7 + [_i] * 4 + [_fp, _sz, _fp, _fp, _fp, _fp]),
    "mmqg_lstm_seq_bwd": (_i, [_fp] * 4 + [_i] * 4 + [_fp, _sz] + [_fp] * 8),
    "mmqg_vocab_workspace_bytes": (_sz, [_i, _i, _i]),
    "mmqg_vocab_nll_fwd": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _f, _fp, _sz, _fp, _fp, _fp, _fp]),
    "mmqg_vocab_nll_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _fp, _sz, _fp, _fp, _fp, _i, _fp]),
    "mmqg_decode_step_argmax": (_i, [_fp, _fp, _fp, _i, _i, _i, _fp, _sz, _fp, _ll, _fp]),
}

_lib = None


class MmqgError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmqgError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                            "there is no CPU or PyTorch fallback for this path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.mmqg_abi_version() != 3:
            raise MmqgError("libmmqg.so ABI version mismatch")
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise MmqgError(f"libmmqg status {status}: {lib().mmqg_last_error().decode()}")


def c_dims(d: Dims) -> MmqgDims:
    return MmqgDims(**{n: getattr(d, n) for n, _ in MmqgDims._fields_})


def ptr(t):
    return 0 if t is None else t.data_ptr()


def c_tensors(p: dict, L: int) -> MmqgTensors:
    """Flat name->CUDA fp32 tensor dict (mmqg.dims.param_shapes names) -> mmqg_tensors."""
    import torch
    for k, v in p.items():
        if not (v.is_cuda and v.dtype == torch.float32 and v.is_contiguous()):
            raise MmqgError(f"{k}: expected a contiguous CUDA float32 tensor, got {v.device} {v.dtype}")
    t = MmqgTensors()
    t.emb = ptr(p["emb.weight"])
    for l in range(L):
        t.text_w_ih[l] = ptr(p[f"text.lstm.weight_ih_l{l}"]); t.text_w_hh[l] = ptr(p[f"text.lstm.weight_hh_l{l}"])
        t.text_b_ih[l] = ptr(p[f"text.lstm.bias_ih_l{l}"]); t.text_b_hh[l] = ptr(p[f"text.lstm.bias_hh_l{l}"])
        t.dec_w_ih[l] = ptr(p[f"dec.lstm.weight_ih_l{l}"]); t.dec_w_hh[l] = ptr(p[f"dec.lstm.weight_hh_l{l}"])
        t.dec_b_ih[l] = ptr(p[f"dec.lstm.bias_ih_l{l}"]); t.dec_b_hh[l] = ptr(p[f"dec.lstm.bias_hh_l{l}"])
    t.vid_w_ih = ptr(p["video.lstm.weight_ih_l0"]); t.vid_w_hh = ptr(p["video.lstm.weight_hh_l0"])
    t.vid_b_ih = ptr(p["video.lstm.bias_ih_l0"]); t.vid_b_hh = ptr(p["video.lstm.bias_hh_l0"])
    for i, n in enumerate(("text_attn", "audio_attn", "vid_attn")):      # 0=text 1=audio 2=video
        t.attn_w[i] = ptr(p[f"dec.{n}.weight"]); t.attn_b[i] = ptr(p[f"dec.{n}.bias"])
    t.out_w = ptr(p["dec.out_layer.weight"]); t.out_b = ptr(p["dec.out_layer.bias"])
    return t
