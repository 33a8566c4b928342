"""Generate the golden fixtures by running the REFERENCE ITSELF (this container only).

    python tests/golden/make_golden.py          # needs /root/reference, writes tests/golden/*.pt

Imports the reference's own model/encoder.py and model/decoder.py, loads synthetic
weights (mmqg.synth.make_params) into them with load_state_dict, and drives them with the
per-sample loop of train.py:153-177 (teacher forcing) and train.py:101-110 (greedy),
exactly as train.py does: one TextEncoder.forward per token, one AttnDecoder.forward per
step, CrossEntropyLoss per step summed, loss.backward().  Stored per case:
  small cases  : weights, inputs, per-sample losses, mean loss, every gradient tensor
                 (mean over samples), per-step logits / attention weights, greedy tokens.
  full-dim case: (E=300, H=512, L=3, TM=283, AM=101; weights are regenerated from the seed,
                 not stored) loss, per-tensor gradient norms, 64 sampled gradient entries
                 per tensor, greedy tokens and top-1/top-2 margins.
The fixtures travel to the GPU box; /root/reference does not.
"""
import contextlib
import io
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))
sys.path.insert(0, "/root/reference")

from mmqg.dims import Dims  # noqa: E402
from mmqg.synth import make_params, make_batch  # noqa: E402
from model.encoder import TextEncoder, VideoConvLstmEncoder  # noqa: E402  (reference)
from model.decoder import AttnDecoder  # noqa: E402  (reference)

CASES = {
    "small_a": dict(dims=Dims(B=3, T_t=6, T_v=3, T_q=5, V=37, E=12, H=32, L=3, H_a=8, H_v=32, F_v=20, TM=9, AM=5),
                    seed=11),
    "small_b": dict(dims=Dims(B=2, T_t=5, T_v=2, T_q=4, V=29, E=10, H=16, L=2, H_a=6, H_v=24, F_v=14, TM=7, AM=4),
                    seed=12),
    "full_dim": dict(dims=Dims(B=2, T_t=12, T_v=4, T_q=5, V=1000, E=300, H=512, L=3, H_a=128, H_v=512, F_v=2048,
                               TM=283, AM=101), seed=13),
}


def build_reference(d, params, dtype):
    dev = torch.device("cpu")
    emb = torch.nn.Embedding(d.V, d.E)
    video = VideoConvLstmEncoder(3, 3, 1, d.H_v, d.F_v)              # only .lstm is on the path
    text = TextEncoder(d.L, 0.0, d.H, d.E, emb, dev)
    dec = AttnDecoder(d.L, 0.0, d.H, d.V, d.E, d.H_v, d.H_a, emb, d.TM, d.AM, dev)
    for m in (emb, video, text, dec):
        m.to(dtype)
    emb.load_state_dict({"weight": params["emb.weight"].to(dtype)})
    video.lstm.load_state_dict({k[len("video.lstm."):]: v.to(dtype) for k, v in params.items()
                                if k.startswith("video.lstm.")})
    text.lstm.load_state_dict({k[len("text.lstm."):]: v.to(dtype) for k, v in params.items()
                               if k.startswith("text.lstm.")})
    sd = {k[len("dec."):]: v.to(dtype) for k, v in params.items() if k.startswith("dec.")}
    sd["emb_layer.weight"] = params["emb.weight"].to(dtype)
    dec.load_state_dict(sd)
    assert text.word_embeddings is emb and dec.emb_layer is emb
    return emb, video, text, dec


def named_grads(emb, video, text, dec):
    out = {"emb.weight": emb.weight.grad}
    for n, p in video.lstm.named_parameters():
        out[f"video.lstm.{n}"] = p.grad
    for n, p in text.lstm.named_parameters():
        out[f"text.lstm.{n}"] = p.grad
    for n, p in dec.named_parameters():
        if not n.startswith("emb_layer"):
            out[f"dec.{n}"] = p.grad
    return {k: v.detach().clone() for k, v in out.items()}


def encode_sample(d, video, text, ctx, frames, audio):
    """train.py:153-166 with features in place of raw media."""
    video_emb = video.lstm(frames.view(frames.shape[0], 1, -1))[0].squeeze(1)   # encoder.py:69,126
    n_frames = video_emb.shape[0]
    padded_audio = F.pad(audio, (0, 0, 0, d.AM - n_frames))
    padded_video = F.pad(video_emb, (0, 0, 0, d.AM - n_frames))
    hid = text.init_state(1)
    hid = tuple(h.to(frames.dtype) for h in hid)
    all_enc = torch.zeros(d.TM, text.hidden_dim, dtype=frames.dtype)
    for ei in range(ctx.shape[0]):
        out, hid = text(ctx[ei], hid)
        all_enc[ei] = out[0, 0]
    return n_frames, padded_audio, padded_video, hid, all_enc


def run_case(name, d, seed, dtype=torch.float64):
    params = make_params(d, seed=seed)
    batch = make_batch(d, seed=seed + 1000)
    emb, video, text, dec = build_reference(d, params, dtype)
    crit = torch.nn.CrossEntropyLoss()
    for m in (emb, video, text, dec):
        m.zero_grad()
    losses, step_logits, step_attn = [], [], []
    with contextlib.redirect_stdout(io.StringIO()):                 # decoder.py:89,97 debug prints
        for b in range(d.B):
            ctx, tgt = batch["context"][b], batch["target"][b]
            fr, au = batch["frames"][b].to(dtype), batch["audio"][b].to(dtype)
            n_frames, pa, pv, hid, all_enc = encode_sample(d, video, text, ctx, fr, au)
            dec_input = torch.tensor([[1]])
            loss = 0
            sl, sa = [], []
            for di in range(d.T_q):                                 # train.py:171-175
                out, hid, a_t, a_a, a_v = dec(dec_input, n_frames, d.T_t, pa, pv, hid, all_enc)
                loss = loss + crit(out, tgt[di].view(-1))
                dec_input = tgt[di]
                sl.append(out.detach()[0])
                sa.append(torch.cat([a_t.detach()[0], a_a.detach()[0], a_v.detach()[0]]))
            (loss / d.B).backward()                                 # mean over samples of the per-sample loss
            losses.append(loss.detach())
            step_logits.append(torch.stack(sl))
            step_attn.append(torch.stack(sa))
    grads = named_grads(emb, video, text, dec)
    fx = {"dims": d.asdict(), "seed": seed, "dtype": str(dtype),
          "loss_per_sample": torch.stack(losses), "loss": torch.stack(losses).mean()}

    # greedy decode on input-sensitive weights: biases x0.1, out weight x10
    gparams = make_params(d, seed=seed, bias_scale=0.1, out_weight_scale=10.0)
    emb, video, text, dec = build_reference(d, gparams, dtype)
    for m in (video, text, dec):
        m.eval()
    max_len = d.T_q + 3
    toks, margins = [], []
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for b in range(d.B):
            fr, au = batch["frames"][b].to(dtype), batch["audio"][b].to(dtype)
            n_frames, pa, pv, hid, all_enc = encode_sample(d, video, text, batch["context"][b], fr, au)
            dec_input = torch.tensor([[1]])
            tk, mg = [], []
            for di in range(max_len):                               # train.py:101-110
                out, hid, *_ = dec(dec_input, n_frames, d.T_t, pa, pv, hid, all_enc)
                word_index = torch.argmax(F.softmax(out, dim=1), dim=1, keepdim=True)
                top2 = out.topk(2, 1).values[0]
                tk.append(int(word_index))
                mg.append(float(top2[0] - top2[1]))
                dec_input = word_index.detach()
            toks.append(tk)
            margins.append(mg)
    fx["greedy_tokens"] = torch.tensor(toks)
    fx["greedy_margins"] = torch.tensor(margins)
    fx["greedy_max_len"] = max_len

    if name.startswith("small"):
        fx["params"] = params
        fx["batch"] = batch
        fx["grads"] = grads
        fx["step_logits"] = torch.stack(step_logits)              # (B,T_q,V)
        fx["step_attn"] = torch.stack(step_attn)                  # (B,T_q,TM+2AM) text|audio|video
    else:
        g = torch.Generator().manual_seed(999)
        fp = {}
        for k, v in grads.items():
            idx = torch.randint(0, v.numel(), (64,), generator=g)
            fp[k] = {"norm": v.norm(), "idx": idx, "vals": v.flatten()[idx].clone()}
        fx["grad_fingerprint"] = fp
    path = os.path.join(HERE, f"{name}.pt")
    torch.save(fx, path)
    print(name, "loss", float(fx["loss"]), "greedy", fx["greedy_tokens"].tolist()[0][:8],
          "min margin", float(fx["greedy_margins"].min()), os.path.getsize(path), "bytes")


VARLEN = {"varlen_a": dict(dims=Dims(B=4, T_t=7, T_v=3, T_q=5, V=41, E=12, H=32, L=3, H_a=8, H_v=32, F_v=20, TM=9, AM=5),
                           seed=21, ctx_len=[7, 1, 4, 5], tgt_len=[5, 2, 1, 4], n_frames=[3, 1, 2, 3])}


def run_varlen_case(name, d, seed, ctx_len, tgt_len, n_frames, dtype=torch.float64):
    """Per-sample lengths: the reference's loop on every sample cut to its own lengths (what a
    batch-1 DataLoader hands train.py); loss = mean over samples, gradients of that mean."""
    params = make_params(d, seed=seed)
    batch = make_batch(d, seed=seed + 1000)
    emb, video, text, dec = build_reference(d, params, dtype)
    crit = torch.nn.CrossEntropyLoss()
    for m in (emb, video, text, dec):
        m.zero_grad()
    losses = []
    with contextlib.redirect_stdout(io.StringIO()):
        for b in range(d.B):
            cl, tl, nf = ctx_len[b], tgt_len[b], n_frames[b]
            ctx, tgt = batch["context"][b, :cl], batch["target"][b, :tl]
            fr, au = batch["frames"][b, :nf].to(dtype), batch["audio"][b, :nf].to(dtype)
            n, pa, pv, hid, all_enc = encode_sample(d, video, text, ctx, fr, au)
            assert n == nf
            dec_input = torch.tensor([[1]])
            loss = 0
            for di in range(tl):
                out, hid, *_ = dec(dec_input, n, cl, pa, pv, hid, all_enc)
                loss = loss + crit(out, tgt[di].view(-1))
                dec_input = tgt[di]
            (loss / d.B).backward()
            losses.append(loss.detach())
    fx = {"dims": d.asdict(), "seed": seed, "dtype": str(dtype), "params": params, "batch": batch,
          "ctx_len": torch.tensor(ctx_len, dtype=torch.int32), "tgt_len": torch.tensor(tgt_len, dtype=torch.int32),
          "n_frames": torch.tensor(n_frames, dtype=torch.int32), "loss_per_sample": torch.stack(losses),
          "loss": torch.stack(losses).mean(), "grads": named_grads(emb, video, text, dec)}
    path = os.path.join(HERE, f"{name}.pt")
    torch.save(fx, path)
    print(name, "loss", float(fx["loss"]), os.path.getsize(path), "bytes")


def run_adam_case(name="adam_a", seed=31, n_iter=3, dtype=torch.float64):
    """Three iterations of the reference's train loop INCLUDING its optimisers, wired exactly as
    train.py:265-267 (Adam over av_enc / text_enc / dec .parameters(), lr 1e-4; the shared embedding
    is a parameter of both the text encoder and the decoder) and stepped as train.py:149-181
    (zero_grad, forward, backward, three .step() per sample).  Stored: the three samples, the
    per-iteration losses and every parameter after the last iteration."""
    d = Dims(B=1, T_t=6, T_v=3, T_q=4, V=37, E=12, H=32, L=3, H_a=8, H_v=32, F_v=20, TM=9, AM=5)
    params = make_params(d, seed=seed)
    emb, video, text, dec = build_reference(d, params, dtype)
    opts = [torch.optim.Adam(video.parameters(), lr=1e-4), torch.optim.Adam(text.parameters(), lr=1e-4),
            torch.optim.Adam(dec.parameters(), lr=1e-4)]
    assert any(p is emb.weight for p in text.parameters()) and any(p is emb.weight for p in dec.parameters())
    crit = torch.nn.CrossEntropyLoss()
    batches, losses = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for it in range(n_iter):
            batch = make_batch(d, seed=seed + 100 + it)
            batches.append(batch)
            for o in opts:
                o.zero_grad()
            fr, au = batch["frames"][0].to(dtype), batch["audio"][0].to(dtype)
            n, pa, pv, hid, all_enc = encode_sample(d, video, text, batch["context"][0], fr, au)
            dec_input = torch.tensor([[1]])
            loss = 0
            for di in range(d.T_q):
                out, hid, *_ = dec(dec_input, n, d.T_t, pa, pv, hid, all_enc)
                loss = loss + crit(out, batch["target"][0][di].view(-1))
                dec_input = batch["target"][0][di]
            loss.backward()
            for o in opts:
                o.step()
            losses.append(float(loss))
    final = {"emb.weight": emb.weight.detach().clone()}
    for n_, p_ in video.lstm.named_parameters():
        final[f"video.lstm.{n_}"] = p_.detach().clone()
    for n_, p_ in text.lstm.named_parameters():
        final[f"text.lstm.{n_}"] = p_.detach().clone()
    for n_, p_ in dec.named_parameters():
        if not n_.startswith("emb_layer"):
            final[f"dec.{n_}"] = p_.detach().clone()
    fx = {"dims": d.asdict(), "seed": seed, "params": params, "batches": batches, "losses": torch.tensor(losses, dtype=torch.float64),
          "final_params": final}
    path = os.path.join(HERE, f"{name}.pt")
    torch.save(fx, path)
    print(name, "losses", losses, os.path.getsize(path), "bytes")


def run_decoder_case(name, T, V, E, A, H, L, seed, dtype=torch.float64):
    """The reference's non-attention Decoder (model/decoder.py:7-47) on one whole question: forward(text (1,T),
    av_enc_out (1,A), hidden) -> logits (T,1,V); loss = sum_t CE(logits_t, target_t) as in non_attn_train.py's
    intent; gradients of every parameter, of av_enc_out and of the initial state."""
    from model.decoder import Decoder                                     # reference
    g = torch.Generator().manual_seed(seed)
    emb = torch.nn.Embedding(V, E)
    dec = Decoder(L, 0.0, H, V, E, A, emb)
    for m in (emb, dec):
        m.to(dtype)
    with torch.no_grad():
        emb.weight.copy_(torch.randn(V, E, generator=g).to(dtype))
        for n_, p_ in dec.named_parameters():
            if n_.startswith("word_embeddings"):
                continue
            if "bias" in n_:
                p_.mul_(0.3)
    text = torch.randint(3, V, (1, T), generator=g)
    target = torch.randint(3, V, (T,), generator=g)
    av = torch.randn(1, A, generator=g).to(dtype).requires_grad_(True)
    h0 = (0.5 * torch.randn(L, 1, H, generator=g)).to(dtype).requires_grad_(True)
    c0 = (0.5 * torch.randn(L, 1, H, generator=g)).to(dtype).requires_grad_(True)
    logits, (hn, cn) = dec(text, av, (h0, c0))
    loss = F.cross_entropy(logits.view(T, V), target, reduction="sum") + 0.1 * hn.sum() + 0.05 * cn.sum()
    loss.backward()
    fx = {"dims": dict(T=T, V=V, E=E, A=A, H=H, L=L), "seed": seed,
          "state_dict": {k: v.detach().float().clone() for k, v in dec.state_dict().items()},
          "text": text, "target": target, "av": av.detach().float(), "h0": h0.detach().float(), "c0": c0.detach().float(),
          "loss": float(loss), "logits": logits.detach().float(), "hn": hn.detach().float(), "cn": cn.detach().float(),
          "grads": {n_: p_.grad.detach().float().clone() for n_, p_ in dec.named_parameters()},
          "d_av": av.grad.detach().float(), "d_h0": h0.grad.detach().float(), "d_c0": c0.grad.detach().float()}
    path = os.path.join(HERE, f"{name}.pt")
    torch.save(fx, path)
    print(name, "loss", float(loss), os.path.getsize(path), "bytes")


def run_convstack_case(name, T, HW, hidden, seed, dtype=torch.float64):
    """f2: the reference's own VideoConvLstmEncoder (model/encoder.py:31-78) on raw frames: conv stack (4 x Conv2d -> ReLU ->
    BatchNorm2d, two MaxPool2d) + LSTM.  One train-mode forward/backward (batch statistics, running buffers updated) and one
    eval-mode forward.  Stored: initial state_dict, frames, outputs, the projection that forms the scalar loss, every
    parameter gradient, the BatchNorm buffers after the train-mode forward."""
    torch.manual_seed(seed)
    s = HW
    for _ in range(2):
        s = ((s - 2) - 2) // 3
    feat = 10 * s * s
    enc = VideoConvLstmEncoder(3, 3, 1, hidden, feat).to(dtype)
    with torch.no_grad():      # BatchNorm affine parameters and running buffers away from their defaults
        for bn in (enc.bn1, enc.bn2, enc.bn3, enc.bn4):
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
            bn.running_mean.uniform_(-0.2, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
        for v in enc.state_dict().values():      # fp32-representable values: the fixture stores fp32 and loses nothing
            if v.is_floating_point():
                v.copy_(v.float().to(dtype))
    state0 = {k: v.detach().clone().float() for k, v in enc.state_dict().items()}
    frames = torch.randn(1, 3, T, HW, HW).to(dtype)
    proj = torch.randn(T, 1, hidden).to(dtype)
    enc.train()
    out = enc(frames)
    loss = (out * proj).sum()
    loss.backward()
    grads = {n_: p_.grad.detach().float().clone() for n_, p_ in enc.named_parameters()}
    state1 = {k: v.detach().clone().float() for k, v in enc.state_dict().items() if "running" in k or "num_batches" in k}
    enc.eval()
    with torch.no_grad():
        out_eval = enc(frames)
    fx = {"T": T, "HW": HW, "hidden": hidden, "feat": feat, "state0": state0, "frames": frames.float(), "proj": proj.float(),
          "out_train": out.detach().float(), "loss": float(loss), "grads": grads, "bn_after": state1, "out_eval": out_eval.float()}
    path = os.path.join(HERE, f"{name}.pt")
    torch.save(fx, path)
    print(name, "loss", float(loss), "features", feat, os.path.getsize(path), "bytes")


CONVSTACK_CASES = {
    "convstack_a": dict(T=3, HW=40, hidden=32, seed=51),       # 40 -> 38 -> 36 -> 12 -> 10 -> 8 -> 2: 40 features
    "convstack_b": dict(T=5, HW=58, hidden=64, seed=52),       # 58 -> ... -> 4: 160 features; hidden on the tensor-core sequence path
}


DECODER_CASES = {
    "decoder_a": dict(T=6, V=37, E=12, A=20, H=32, L=2, seed=41),      # fp32 building blocks (H not a multiple of 64)
    "decoder_b": dict(T=9, V=53, E=12, A=20, H=64, L=3, seed=42),      # shape the tensor-core sequence kernels take
}


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1:]
    for name, c in CONVSTACK_CASES.items():
        if not which or name in which:
            run_convstack_case(name, **c)
    for name, c in DECODER_CASES.items():
        if not which or name in which:
            run_decoder_case(name, **c)
    if which and all(w in DECODER_CASES or w in CONVSTACK_CASES for w in which):
        sys.exit(0)
    if not which or "adam_a" in which:
        run_adam_case()
    for name, c in CASES.items():
        if not which or name in which:
            run_case(name, c["dims"], c["seed"])
    for name, c in VARLEN.items():
        if not which or name in which:
            run_varlen_case(name, c["dims"], c["seed"], c["ctx_len"], c["tgt_len"], c["n_frames"])
