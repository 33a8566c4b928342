"""Drop-in boundary: the `model` package of this repo mirrors the reference's model/encoder.py and
model/decoder.py (class names, constructor / forward signatures, state_dict keys), and the
reference's per-sample training loop runs on it unchanged in structure.

CPU part: signature / state_dict parity against /root/reference when it is present (this
container) and against a recorded copy of the contract otherwise.  GPU part: the loop of
train.py:149-177 (restated in this file, since /root/reference does not travel to the GPU box)
drives the modules sample by sample; losses and gradients must match the golden fixtures that
the reference itself produced."""
import importlib.util
import inspect
import os

import pytest
import torch

from conftest import load_golden
from mmqg.dims import Dims

REF = "/root/reference"

# the contract, recorded from the reference
CTOR_ARGS = {
    "TextEncoder": ["num_layers", "dropout_p", "hidden_dim", "emb_dim", "emb_layer", "device"],
    "VideoConvLstmEncoder": ["in_channels", "kernel_sz", "stride", "hidden_dim", "video_emb_dim"],
    "AudioVideoEncoder": ["av_in_channels", "av_kernel_sz", "av_stride", "av_hidden_dim", "video_emb_dim"],
    "AttnDecoder": ["num_layers", "dropout_p", "hidden_dim", "n_vocab", "word_emb_dim", "video_emb_dim", "audio_emb_dim",
                    "emb_layer", "text_max_length", "av_max_length", "device"],
    "Decoder": ["num_layers", "dropout", "hidden_dim", "n_vocab", "word_emb_dim", "av_emb_dim", "emb_layer"],
}
FORWARD_ARGS = {
    "TextEncoder": ["text", "hidden"],
    "VideoConvLstmEncoder": ["video_frames"],
    "AudioVideoEncoder": ["audio_file", "video_frames"],
    "AttnDecoder": ["word", "enc_frames", "enc_seq_len", "audio_emb", "video_emb", "hidden", "encoder_outputs"],
    "Decoder": ["text", "av_enc_out", "hidden"],
}


def _ours():
    import model.encoder as E
    import model.decoder as D
    assert "multi-modal-qg_b200" in E.__file__
    return E, D


def _args(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def test_signatures_match_contract():
    E, D = _ours()
    for name, want in CTOR_ARGS.items():
        cls = getattr(E, name, None) or getattr(D, name)
        assert _args(cls.__init__) == want, name
        assert _args(cls.forward) == FORWARD_ARGS[name], name
    assert _args(E.TextEncoder.init_state) == ["batch_sz"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_signatures_and_state_dict_match_reference():
    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    RE = load("ref_encoder", os.path.join(REF, "model", "encoder.py"))
    RD = load("ref_decoder", os.path.join(REF, "model", "decoder.py"))
    E, D = _ours()
    for name in CTOR_ARGS:
        ours = getattr(E, name, None) or getattr(D, name)
        ref = getattr(RE, name, None) or getattr(RD, name)
        assert _args(ours.__init__) == _args(ref.__init__), name
        assert _args(ours.forward) == _args(ref.forward), name
    emb = torch.nn.Embedding(50, 12)
    dev = torch.device("cpu")
    pairs = [
        (E.TextEncoder(3, 0.2, 32, 12, emb, dev), RE.TextEncoder(3, 0.2, 32, 12, emb, dev)),
        (E.VideoConvLstmEncoder(3, 3, 1, 32, 1000), RE.VideoConvLstmEncoder(3, 3, 1, 32, 1000)),
        (D.AttnDecoder(3, 0.2, 32, 50, 12, 32, 8, emb, 9, 5, dev), RD.AttnDecoder(3, 0.2, 32, 50, 12, 32, 8, emb, 9, 5, dev)),
        (D.Decoder(2, 0.1, 32, 50, 12, 16, emb), RD.Decoder(2, 0.1, 32, 50, 12, 16, emb)),
    ]
    for ours, ref in pairs:
        so, sr = ours.state_dict(), ref.state_dict()
        assert list(so) == list(sr), type(ours).__name__
        for k in so:
            assert so[k].shape == sr[k].shape and so[k].dtype == sr[k].dtype, k
        ours.load_state_dict(sr, strict=True)          # checkpoints round-trip both ways
        ref.load_state_dict(so, strict=True)
    # same initialiser sequence as the reference: identical weights under the same seed
    torch.manual_seed(5)
    a = E.TextEncoder(2, 0.0, 16, 8, torch.nn.Embedding(9, 8), dev)
    torch.manual_seed(5)
    b = RE.TextEncoder(2, 0.0, 16, 8, torch.nn.Embedding(9, 8), dev)
    for (k, v), (_, u) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, u), k
    # shared embedding identity is preserved (train.py:236,245,255)
    t, dcd = pairs[0][0], pairs[2][0]
    assert t.word_embeddings is emb and dcd.emb_layer is emb


def test_modules_refuse_cpu_tensors():
    E, D = _ours()
    emb = torch.nn.Embedding(20, 8)
    enc = E.TextEncoder(2, 0.0, 16, 8, emb, torch.device("cpu"))
    with pytest.raises(RuntimeError, match="CUDA only"):
        enc(torch.tensor(3), enc.init_state(1))


# ---- GPU: the reference's per-sample loop on the drop-in modules ----------------------------------
def _build_modules(d, params, dev):
    E, D = _ours()
    emb = torch.nn.Embedding(d.V, d.E)
    video = E.VideoConvLstmEncoder(3, 3, 1, d.H_v, d.F_v)
    text = E.TextEncoder(d.L, 0.0, d.H, d.E, emb, dev)
    dec = D.AttnDecoder(d.L, 0.0, d.H, d.V, d.E, d.H_v, d.H_a, emb, d.TM, d.AM, dev)
    emb.load_state_dict({"weight": params["emb.weight"].float()})
    video.lstm.load_state_dict({k[len("video.lstm."):]: v.float() for k, v in params.items() if k.startswith("video.lstm.")})
    text.lstm.load_state_dict({k[len("text.lstm."):]: v.float() for k, v in params.items() if k.startswith("text.lstm.")})
    sd = {k[len("dec."):]: v.float() for k, v in params.items() if k.startswith("dec.")}
    sd["emb_layer.weight"] = params["emb.weight"].float()
    dec.load_state_dict(sd)
    for m in (video, text, dec):
        m.to(dev)
    return emb, video, text, dec


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small_a", "small_b"])
def test_reference_loop_on_dropin_modules(name):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import torch.nn.functional as F
    fx = load_golden(name)
    d = Dims(**fx["dims"])
    dev = torch.device("cuda")
    emb, video, text, dec = _build_modules(d, fx["params"], dev)
    criterion = torch.nn.CrossEntropyLoss()
    batch = fx["batch"]
    for m in (video, text, dec):
        m.train()
        m.zero_grad()
    losses = []
    for b in range(d.B):                                                     # train.py:144 (batch_size=1)
        ctx = batch["context"][b:b + 1].to(dev)
        target = batch["target"][b:b + 1].to(dev)
        frames, audio = batch["frames"][b].to(dev), batch["audio"][b].to(dev)
        video_emb = video.encode_features(frames).squeeze(1)                 # feature-level stand-in for :153
        n_frames = video_emb.shape[0]
        padded_audio_emb = F.pad(audio, (0, 0, 0, d.AM - n_frames))          # :156
        padded_video_emb = F.pad(video_emb, (0, 0, 0, d.AM - n_frames))      # :157
        text_enc_hidden = text.init_state(1)                                 # :159
        all_enc_outputs = torch.zeros(d.TM, text.hidden_dim).to(dev)         # :160
        loss = 0
        for ei in range(d.T_t):                                              # :164-166
            enc_output, text_enc_hidden = text(ctx[0][ei], text_enc_hidden)
            all_enc_outputs[ei] = enc_output[0, 0]
        dec_input = torch.tensor([[1]]).to(dev)                              # :168
        dec_hidden = text_enc_hidden                                         # :169
        for di in range(d.T_q):                                              # :171-175
            dec_output, dec_hidden, text_attn, audio_attn, vid_attn = dec(
                dec_input, n_frames, d.T_t, padded_audio_emb, padded_video_emb, dec_hidden, all_enc_outputs)
            loss += criterion(dec_output, target[0][di].view(-1))
            dec_input = target[0][di]
        (loss / d.B).backward()                                              # :177 (mean over the batch)
        losses.append(float(loss))
        assert text_attn.shape == (1, d.TM) and audio_attn.shape == (1, d.AM) and vid_attn.shape == (1, d.AM)
    want = fx["loss_per_sample"].double()
    got = torch.tensor(losses, dtype=torch.float64)
    assert torch.allclose(got, want, rtol=1e-4), (got, want)
    grads = {"emb.weight": emb.weight.grad}
    grads.update({f"video.lstm.{n}": p.grad for n, p in video.lstm.named_parameters()})
    grads.update({f"text.lstm.{n}": p.grad for n, p in text.lstm.named_parameters()})
    grads.update({f"dec.{n}": p.grad for n, p in dec.named_parameters() if not n.startswith("emb_layer")})
    for k, g in fx["grads"].items():
        a, r = grads[k].double().cpu(), g.double()
        assert float((a - r).norm() / r.norm().clamp_min(1e-30)) < 1e-3, k


@pytest.mark.gpu
def test_dropin_greedy_matches_golden():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import torch.nn.functional as F
    from mmqg.synth import make_batch, make_params
    fx = load_golden("small_a")
    d = Dims(**fx["dims"])
    dev = torch.device("cuda")
    gp = make_params(d, seed=fx["seed"], bias_scale=0.1, out_weight_scale=10.0)
    batch = make_batch(d, seed=fx["seed"] + 1000)
    emb, video, text, dec = _build_modules(d, gp, dev)
    for m in (video, text, dec):
        m.eval()
    toks = []
    with torch.no_grad():
        for b in range(d.B):                                                 # train.py:76-110
            frames, audio = batch["frames"][b].to(dev), batch["audio"][b].to(dev)
            video_emb = video.encode_features(frames).squeeze(1)
            n_frames = video_emb.shape[0]
            pa = F.pad(audio, (0, 0, 0, d.AM - n_frames))
            pv = F.pad(video_emb, (0, 0, 0, d.AM - n_frames))
            hid = text.init_state(1)
            all_enc = torch.zeros(d.TM, text.hidden_dim).to(dev)
            for ei in range(d.T_t):
                out, hid = text(batch["context"][b][ei].to(dev), hid)
                all_enc[ei] = out[0, 0]
            dec_input = torch.tensor([[1]]).to(dev)
            row = []
            for di in range(fx["greedy_max_len"]):
                out, hid, *_ = dec(dec_input, n_frames, d.T_t, pa, pv, hid, all_enc)
                word_index = torch.argmax(F.softmax(out, dim=1), dim=1, keepdim=True)
                row.append(int(word_index))
                dec_input = word_index.detach()
            toks.append(row)
    assert toks == fx["greedy_tokens"].tolist()
