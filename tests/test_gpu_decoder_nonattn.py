"""the non-attention `Decoder` (reference model/decoder.py:7-47) as a drop-in, checked against
fixtures produced by the reference's OWN class (tests/golden/make_golden.py run_decoder_case; fp64).

fp32 drop-in mode (building blocks): 1e-3 relative on loss, logits, state and every gradient (north_star bar).
bf16 drop-in mode (MMQG_DROPIN_MODE=bf16: the whole-sequence LSTM call runs on mmqg_lstm_seq_fwd / _bwd, i.e. the
hoisted tcgen05 input projection + persistent recurrent / BPTT kernels): logits 2e-2, gradients 5e-2 relative --
the bf16 bar of tests/test_gpu_bf16_mode.py; weights, inputs and h are rounded to bf16 inside."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_dropin(fx, monkeypatch, mode):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    monkeypatch.setenv("MMQG_DROPIN_MODE", mode)
    from model.decoder import Decoder            # the drop-in (multi-modal-qg_b200/model on sys.path via conftest)
    import torch.nn.functional as F
    d = fx["dims"]
    dev = torch.device("cuda")
    emb = torch.nn.Embedding(d["V"], d["E"])
    dec = Decoder(d["L"], 0.0, d["H"], d["V"], d["E"], d["A"], emb)
    dec.load_state_dict(fx["state_dict"], strict=True)
    dec.to(dev)
    av = fx["av"].to(dev).requires_grad_(True)
    h0 = fx["h0"].to(dev).requires_grad_(True)
    c0 = fx["c0"].to(dev).requires_grad_(True)
    logits, (hn, cn) = dec(fx["text"].to(dev), av, (h0, c0))
    T, V = d["T"], d["V"]
    loss = F.cross_entropy(logits.view(T, V), fx["target"].to(dev), reduction="sum") + 0.1 * hn.sum() + 0.05 * cn.sum()
    loss.backward()
    torch.cuda.synchronize()
    out = {"loss": abs(float(loss) - fx["loss"]) / abs(fx["loss"]), "logits": rel(logits, fx["logits"]),
           "hn": rel(hn, fx["hn"]), "cn": rel(cn, fx["cn"]), "d_av": rel(av.grad, fx["d_av"]),
           "d_h0": rel(h0.grad, fx["d_h0"]), "d_c0": rel(c0.grad, fx["d_c0"])}
    for n, p in dec.named_parameters():
        out["grad " + n] = rel(p.grad, fx["grads"][n])
    return out


@pytest.mark.parametrize("name", ["decoder_a", "decoder_b"])
def test_decoder_fp32_matches_reference_fixture(name, monkeypatch):
    errs = run_dropin(load_golden(name), monkeypatch, "fp32")
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"{name} fp32: worst {worst}")
    assert worst[1] < 1e-3, errs


def test_decoder_bf16_sequence_kernels_match_reference_fixture(monkeypatch):
    from mmqg import _cabi
    n0 = int(_cabi.lib().mmqg_launch_count()) if torch.cuda.is_available() else 0
    errs = run_dropin(load_golden("decoder_b"), monkeypatch, "bf16")
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"decoder_b bf16: worst {worst}; loss {errs['loss']:.2e} logits {errs['logits']:.2e}")
    assert errs["loss"] < 5e-3 and errs["logits"] < 2e-2 and errs["hn"] < 2e-2 and errs["cn"] < 2e-2, errs
    assert worst[1] < 5e-2, errs
    assert int(_cabi.lib().mmqg_launch_count()) > n0
