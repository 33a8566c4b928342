"""Per-sample CPU port of the reference's execution structure.  TEST INFRASTRUCTURE ONLY.

The reference runs batch-size-1 (train.py:233-234): for every sample it calls stock
torch.nn.LSTM / Linear / Embedding modules once per token (train.py:164-166) and once
per decoder step (train.py:171-175), sums the per-step CrossEntropyLoss and calls
loss.backward() (train.py:177).  This module is a port of that loop (not of its source
text): the same stock torch.nn modules, the same call granularity, the same order,
fed with features instead of raw media (AudioEncoder needs the network).
bench.py times it as the CPU baseline ("port"); tests use it as a second, independent
statement of the per-sample semantics next to oracle/mmqg_oracle.py.
"""
import torch
from torch import nn
import torch.nn.functional as F

START, END = 1, 2


class RefModules:
    """Stock torch.nn modules holding the parameters of a flat param dict."""

    def __init__(self, params, L, dropout_p=0.0, dtype=torch.float32):
        V, E = params["emb.weight"].shape
        H = params["text.lstm.weight_hh_l0"].shape[1]
        H_v = params["video.lstm.weight_hh_l0"].shape[1]
        F_v = params["video.lstm.weight_ih_l0"].shape[1]
        X0 = params["dec.lstm.weight_ih_l0"].shape[1]
        TM, Q = params["dec.text_attn.weight"].shape
        AM = params["dec.vid_attn.weight"].shape[0]
        self.L, self.H, self.TM, self.AM = L, H, TM, AM
        self.emb = nn.Embedding(V, E)                       # shared (train.py:236,245,255)
        self.text_lstm = nn.LSTM(E, H, L, dropout=dropout_p)
        self.video_lstm = nn.LSTM(F_v, H_v)
        self.text_attn = nn.Linear(Q, TM)
        self.vid_attn = nn.Linear(Q, AM)
        self.audio_attn = nn.Linear(Q, AM)
        self.dec_lstm = nn.LSTM(X0, H, L, dropout=dropout_p)
        self.out_layer = nn.Linear(H, V)
        self._named = {"emb": self.emb, "text.lstm": self.text_lstm, "video.lstm": self.video_lstm,
                       "dec.text_attn": self.text_attn, "dec.vid_attn": self.vid_attn,
                       "dec.audio_attn": self.audio_attn, "dec.lstm": self.dec_lstm,
                       "dec.out_layer": self.out_layer}
        for prefix, m in self._named.items():
            m.to(dtype)
            sd = {k[len(prefix) + 1:]: v.to(dtype) for k, v in params.items() if k.startswith(prefix + ".")}
            m.load_state_dict(sd, strict=True)

    def modules(self):
        return list(self._named.values())

    def named_parameters(self):
        for prefix, m in self._named.items():
            for n, p in m.named_parameters():
                yield f"{prefix}.{n}", p

    def zero_grad(self):
        for m in self.modules():
            m.zero_grad(set_to_none=True)

    def train(self, mode=True):
        for m in self.modules():
            m.train(mode)

    # one sample, exactly the call granularity of train.py:153-175 --------------------
    def encode_sample(self, ctx, frames, audio):
        H, L = self.H, self.L
        vid = self.video_lstm(frames.unsqueeze(1))[0].squeeze(1)             # encoder.py:69,126
        n = vid.shape[0]
        M_aud = F.pad(audio, (0, 0, 0, self.AM - n))                         # train.py:156
        M_vid = F.pad(vid, (0, 0, 0, self.AM - n))                           # train.py:157
        hid = (torch.zeros(L, 1, H, dtype=vid.dtype), torch.zeros(L, 1, H, dtype=vid.dtype))
        M_txt = torch.zeros(self.TM, H, dtype=vid.dtype)                     # train.py:160
        for ei in range(ctx.shape[0]):                                       # train.py:164-166
            e = self.emb(ctx[ei].view(1, -1))
            out, hid = self.text_lstm(e.view(1, 1, -1), hid)
            M_txt[ei] = out[0, 0]
        return M_txt, M_aud, M_vid, hid

    def decoder_step(self, word, hid, M_txt, M_aud, M_vid):                  # decoder.py:74-107
        e = self.emb(word).view(1, 1, -1)
        q = torch.cat((e[0], hid[0][-1]), 1)
        a_txt = F.softmax(self.text_attn(q), dim=1)
        a_vid = F.softmax(self.vid_attn(q), dim=1)
        a_aud = F.softmax(self.audio_attn(q), dim=1)
        c_txt = torch.bmm(a_txt.unsqueeze(0), M_txt.unsqueeze(0))
        c_vid = torch.bmm(a_vid.unsqueeze(0), M_vid.unsqueeze(0))
        c_aud = torch.bmm(a_aud.unsqueeze(0), M_aud.unsqueeze(0))
        x = torch.cat((e[0], c_txt[0], c_aud[0], c_vid[0]), 1).unsqueeze(0)
        out, hid = self.dec_lstm(x, hid)
        return self.out_layer(out[0]), hid, a_txt, a_aud, a_vid

    def sample_loss(self, ctx, frames, audio, tgt):
        M_txt, M_aud, M_vid, hid = self.encode_sample(ctx, frames, audio)
        word = torch.tensor([[START]])
        loss = 0
        for di in range(tgt.shape[0]):
            logits, hid, *_ = self.decoder_step(word, hid, M_txt, M_aud, M_vid)
            loss = loss + F.cross_entropy(logits, tgt[di].view(-1))           # train.py:174
            word = tgt[di]
        return loss

    @torch.no_grad()
    def sample_greedy(self, ctx, frames, audio, max_len):
        M_txt, M_aud, M_vid, hid = self.encode_sample(ctx, frames, audio)
        word = torch.tensor([[START]])
        toks = []
        for _ in range(max_len):
            logits, hid, *_ = self.decoder_step(word, hid, M_txt, M_aud, M_vid)
            word = torch.argmax(F.softmax(logits, dim=1), dim=1, keepdim=True)  # train.py:107-108
            toks.append(int(word))
        return toks


def train_samples(ref: RefModules, batch, n_samples=None, optimizers=None):
    """Run the reference's per-sample train iteration (zero_grad, fwd, bwd[, Adam]) over the
    first n_samples of the batch.  Returns the per-sample losses.  Gradients of the LAST
    sample stay in .grad (as in the reference, which zeroes them every iteration)."""
    n = batch["context"].shape[0] if n_samples is None else n_samples
    losses = []
    for b in range(n):
        ref.zero_grad()
        loss = ref.sample_loss(batch["context"][b], batch["frames"][b], batch["audio"][b], batch["target"][b])
        loss.backward()
        if optimizers:
            for o in optimizers:
                o.step()
        losses.append(float(loss))
    return losses


def batch_loss_and_grads(params, batch, L, dtype=torch.float64):
    """Mean over samples of the per-sample loss and its gradient: the batched quantity
    the CUDA path computes, obtained the slow way (per-sample loop + accumulation)."""
    ref = RefModules(params, L, 0.0, dtype)
    B = batch["context"].shape[0]
    total = 0.0
    ref.zero_grad()
    for b in range(B):
        fr = batch["frames"][b].to(dtype)
        au = batch["audio"][b].to(dtype)
        loss = ref.sample_loss(batch["context"][b], fr, au, batch["target"][b]) / B
        loss.backward()
        total += float(loss.detach())
    grads = {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for n, p in ref.named_parameters()}
    return total, grads
