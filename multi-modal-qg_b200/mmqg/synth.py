"""Synthetic weights and batches of the BASELINE.json shapes.

There is no network for GloVe / VGGish / the dataset, so every run uses random-init
weights with the reference's initialisers (orthogonal LSTM weights, N(0,1) biases,
xavier Linear weights: reference model/encoder.py:102-107, model/decoder.py:109-125;
N(0,1) embedding standing in for GloVe, train.py:25-31) and random inputs of the
named shapes.  Everything is generated on the CPU from explicit generators, so the
same call gives the same tensors in this container and on the GPU box.
"""
import torch
from torch.nn import init

from .dims import Dims, param_shapes

PAD, START, END = 0, 1, 2      # reference prepare_data.py:63-66


def make_params(d: Dims, seed: int = 0, dtype=torch.float32, bias_scale: float = 1.0,
                out_weight_scale: float = 1.0) -> dict:
    """name -> CPU tensor.  bias_scale / out_weight_scale give the input-sensitive
    weights the greedy-parity test needs (last finding)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in param_shapes(d).items():
        t = torch.empty(shape, dtype=torch.float32)
        if name == "emb.weight":
            init.normal_(t, generator=g)
        elif ".lstm.weight" in name:
            init.orthogonal_(t, generator=g)
        elif "bias" in name:
            init.normal_(t, generator=g)
            t.mul_(bias_scale)
        else:                                   # Linear weights
            init.xavier_uniform_(t, generator=g)
            if name == "dec.out_layer.weight":
                t.mul_(out_weight_scale)
        p[name] = t.to(dtype)
    return p


def make_batch(d: Dims, seed: int = 1234) -> dict:
    """Uniform-length synthetic batch.  Keys mirror the tuple VQGDataset yields
    (reference utils/dataset.py:55) with features in place of raw media."""
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randint(3, d.V, (d.B, d.T_t), generator=g, dtype=torch.int64)
    tgt = torch.randint(3, d.V, (d.B, d.T_q), generator=g, dtype=torch.int64)
    tgt[:, -1] = END
    frames = torch.randn(d.B, d.T_v, d.F_v, generator=g)
    audio = torch.relu(torch.randn(d.B, d.T_v, d.H_a, generator=g))
    return {"context": ctx, "target": tgt, "frames": frames, "audio": audio}


def round_params_bf16(p: dict) -> dict:
    """Weights rounded once to bf16 and widened back: what the bf16 mode's packed
    weight caches hold.  Used to build the same-rounded-weights oracle twin."""
    return {k: v.to(torch.bfloat16).to(v.dtype) for k, v in p.items()}
