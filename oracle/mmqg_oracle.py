"""CPU restatement of the multi-modal-qg training hot path.  TEST INFRASTRUCTURE ONLY.

Plain tensor arithmetic (matmul / sigmoid / tanh / exp) on the CPU, batched over the
sample dimension, dtype-generic (fp32 or fp64).  No torch.nn modules: every formula
is written out so that it can be compared line by line with the CUDA kernels.
Gradients come from autograd over this arithmetic.

Reference lines restated (paths relative to /root/reference):
  * LSTM cell, gate order i,f,g,o, both biases  -> torch.nn.LSTM as configured at
    model/encoder.py:91,54 and model/decoder.py:69 (the arithmetic itself lives in
    PyTorch/oneDNN, a dependency the reference does not pin).
  * text encoder                                -> model/encoder.py:95-100, train.py:159-166
  * video LSTM + zero padding                   -> model/encoder.py:69, train.py:155-157
  * attention decoder step                      -> model/decoder.py:74-107
  * teacher-forced loss (sum over steps of batch-mean CE) -> train.py:168-175,264
  * greedy decode                               -> train.py:101-110, evaluate.py:70,101-103

The batched semantics are "the reference's per-sample loop applied to every sample";
tests/test_oracle_golden.py checks exactly that against fixtures produced by the
reference itself (tests/golden/make_golden.py).
"""
import torch

START, END = 1, 2


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    g = x @ w_ih.t() + h @ w_hh.t() + b_ih + b_hh
    H = h.shape[-1]
    i, f, gg, o = g[..., :H], g[..., H:2 * H], g[..., 2 * H:3 * H], g[..., 3 * H:]
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def _layer_weights(p, prefix, l):
    return (p[f"{prefix}.weight_ih_l{l}"], p[f"{prefix}.weight_hh_l{l}"],
            p[f"{prefix}.bias_ih_l{l}"], p[f"{prefix}.bias_hh_l{l}"])


def lstm_stack_step(p, prefix, L, x, h, c, drop_masks=None):
    """One timestep through L stacked cells.  h, c: lists of (B,H).  drop_masks: list of
    L-1 multiplicative masks (already scaled by 1/(1-p)) applied to the outputs of layers
    0..L-2, or None (eval / p=0) -- torch.nn.LSTM's inter-layer dropout."""
    hn, cn = [], []
    for l in range(L):
        h2, c2 = lstm_cell(x, h[l], c[l], *_layer_weights(p, prefix, l))
        hn.append(h2)
        cn.append(c2)
        x = h2
        if drop_masks is not None and l < L - 1:
            x = x * drop_masks[l]
    return x, hn, cn


def video_encode(p, frames, AM):
    """frames (B,T_v,F_v) -> zero-padded video memory (B,AM,H_v)."""
    B, T_v, _ = frames.shape
    H_v = p["video.lstm.weight_hh_l0"].shape[1]
    h = frames.new_zeros(B, H_v)
    c = frames.new_zeros(B, H_v)
    rows = []
    for t in range(T_v):
        h, c = lstm_cell(frames[:, t], h, c, *_layer_weights(p, "video.lstm", 0))
        rows.append(h)
    mem = torch.stack(rows, 1)
    return torch.nn.functional.pad(mem, (0, 0, 0, AM - T_v))


def audio_memory(audio, AM):
    """audio (B,T_a,H_a) -> zero-padded (B,AM,H_a)  (identity + pad, train.py:156)."""
    return torch.nn.functional.pad(audio, (0, 0, 0, AM - audio.shape[1]))


def text_encode(p, ctx, L, TM, drop_masks=None):
    """ctx (B,T_t) int64 -> text memory (B,TM,H) (rows >= T_t zero) and final (h,c) lists."""
    B, T_t = ctx.shape
    H = p["text.lstm.weight_hh_l0"].shape[1]
    emb = p["emb.weight"]
    h = [emb.new_zeros(B, H) for _ in range(L)]
    c = [emb.new_zeros(B, H) for _ in range(L)]
    rows = []
    for t in range(T_t):
        x = emb[ctx[:, t]]
        m = None if drop_masks is None else drop_masks[t]
        top, h, c = lstm_stack_step(p, "text.lstm", L, x, h, c, m)
        rows.append(top)
    mem = torch.stack(rows, 1)
    return torch.nn.functional.pad(mem, (0, 0, 0, TM - T_t)), h, c


def decoder_step(p, L, word, h, c, M_txt, M_aud, M_vid, drop_masks=None):
    """One AttnDecoder.forward (decoder.py:74-107), batched.  Returns logits (B,V), new
    h, c lists and the three attention weight matrices (text, audio, video)."""
    e = p["emb.weight"][word]                                   # decoder.py:75
    q = torch.cat([e, h[L - 1]], 1)                             # decoder.py:78
    # No length mask: decoder.py:79,85,93 index dim 0 of a (1,L) tensor -> no-op (Q1).
    a_txt = torch.softmax(q @ p["dec.text_attn.weight"].t() + p["dec.text_attn.bias"], 1)
    a_vid = torch.softmax(q @ p["dec.vid_attn.weight"].t() + p["dec.vid_attn.bias"], 1)
    a_aud = torch.softmax(q @ p["dec.audio_attn.weight"].t() + p["dec.audio_attn.bias"], 1)
    c_txt = torch.bmm(a_txt.unsqueeze(1), M_txt).squeeze(1)     # decoder.py:81
    c_vid = torch.bmm(a_vid.unsqueeze(1), M_vid).squeeze(1)     # decoder.py:87
    c_aud = torch.bmm(a_aud.unsqueeze(1), M_aud).squeeze(1)     # decoder.py:95
    x = torch.cat([e, c_txt, c_aud, c_vid], 1)                  # decoder.py:99
    top, h, c = lstm_stack_step(p, "dec.lstm", L, x, h, c, drop_masks)
    logits = top @ p["dec.out_layer.weight"].t() + p["dec.out_layer.bias"]   # decoder.py:106
    return logits, h, c, a_txt, a_aud, a_vid


def encode(p, batch, L, TM, AM, text_drop=None):
    """text_drop: multiplicative inter-layer dropout masks (L-1, T_t, B, H) or None."""
    M_vid = video_encode(p, batch["frames"], AM)
    M_aud = audio_memory(batch["audio"], AM)
    masks = None if text_drop is None else [[text_drop[l, t] for l in range(L - 1)] for t in range(text_drop.shape[1])]
    M_txt, h, c = text_encode(p, batch["context"], L, TM, masks)
    return M_txt, M_aud, M_vid, h, c


def teacher_forced_loss(p, batch, L, TM, AM, return_steps=False, drop=None):
    """loss = sum_t mean_b NLL(b,t)  (train.py:171-175 with CrossEntropyLoss() mean).
    drop: None (eval / p=0) or {"text": (L-1,T_t,B,H), "dec": (L-1,T_q,B,H)} multiplicative
    masks (0 or 1/(1-p)) of torch.nn.LSTM's inter-layer dropout (encoder.py:62, decoder.py:57)."""
    M_txt, M_aud, M_vid, h, c = encode(p, batch, L, TM, AM, None if drop is None else drop["text"])
    tgt = batch["target"]
    B, T_q = tgt.shape
    word = torch.full((B,), START, dtype=torch.int64)
    loss = 0
    steps = []
    for t in range(T_q):
        dm = None if drop is None else [drop["dec"][l, t] for l in range(L - 1)]
        logits, h, c, a_txt, a_aud, a_vid = decoder_step(p, L, word, h, c, M_txt, M_aud, M_vid, dm)
        lse = torch.logsumexp(logits, 1)
        nll = lse - logits.gather(1, tgt[:, t:t + 1]).squeeze(1)
        loss = loss + nll.mean()
        word = tgt[:, t]                                        # teacher forcing, train.py:175
        if return_steps:
            steps.append({"logits": logits.detach(), "nll": nll.detach(), "a_txt": a_txt.detach(),
                          "a_aud": a_aud.detach(), "a_vid": a_vid.detach(), "h_top": h[L - 1].detach()})
    if return_steps:
        return loss, {"M_txt": M_txt.detach(), "M_vid": M_vid.detach(), "steps": steps}
    return loss


def teacher_forced_loss_varlen(p, batch, L, TM, AM):
    """Batch with per-sample lengths (batch["ctx_len"], ["tgt_len"], ["n_frames"], each (B) or absent):
    literally the reference's per-sample loop (train.py:153-177) on every sample cut to its own
    lengths; loss = mean over samples of the per-sample summed cross entropy."""
    B = batch["context"].shape[0]
    total = 0
    for b in range(B):
        cl = int(batch["ctx_len"][b]) if "ctx_len" in batch else batch["context"].shape[1]
        tl = int(batch["tgt_len"][b]) if "tgt_len" in batch else batch["target"].shape[1]
        nf = int(batch["n_frames"][b]) if "n_frames" in batch else batch["frames"].shape[1]
        one = {"context": batch["context"][b:b + 1, :cl], "target": batch["target"][b:b + 1, :tl],
               "frames": batch["frames"][b:b + 1, :nf], "audio": batch["audio"][b:b + 1, :nf]}
        total = total + teacher_forced_loss(p, one, L, TM, AM)
    return total / B


def loss_and_grads(params, batch, L, TM, AM, dtype=torch.float64, drop=None):
    """Loss and d loss / d every parameter, computed in `dtype` on the CPU."""
    p = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    b = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}
    if drop is not None:
        drop = {k: v.to(dtype) for k, v in drop.items()}
    if any(k in b for k in ("ctx_len", "tgt_len", "n_frames")):
        assert drop is None, "variable lengths and dropout masks are not combined in the oracle"
        loss = teacher_forced_loss_varlen(p, b, L, TM, AM)
    else:
        loss = teacher_forced_loss(p, b, L, TM, AM, drop=drop)
    names = list(p)
    grads = torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)
    g = {n: (torch.zeros_like(p[n]) if gr is None else gr).detach() for n, gr in zip(names, grads)}
    return loss.detach(), g


@torch.no_grad()
def greedy_decode(params, batch, L, TM, AM, max_len, dtype=torch.float64, return_margins=False):
    """Greedy tokens (B,max_len): argmax of the logits, first index on ties
    (train.py:107-108); no early exit here -- callers cut at <end> (evaluate.py:101-103)."""
    p = {k: v.detach().to(dtype) for k, v in params.items()}
    b = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in batch.items()}
    M_txt, M_aud, M_vid, h, c = encode(p, b, L, TM, AM)
    B = b["context"].shape[0]
    word = torch.full((B,), START, dtype=torch.int64)
    toks, margins = [], []
    for _ in range(max_len):
        logits, h, c, *_ = decoder_step(p, L, word, h, c, M_txt, M_aud, M_vid)
        word = torch.argmax(logits, 1)
        toks.append(word)
        if return_margins:
            top2 = logits.topk(2, 1).values
            margins.append(top2[:, 0] - top2[:, 1])
    toks = torch.stack(toks, 1)
    if return_margins:
        return toks, torch.stack(margins, 1)
    return toks


def rel_err(a, b):
    """Per-tensor relative error ||a-b|| / ||b|| used for the 1e-3 bar (north_star)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
