"""The oracle is pinned against fixtures produced by the reference itself
(tests/golden/make_golden.py): loss, every gradient, per-step logits / attention
weights, greedy tokens.  CPU only."""
import pytest
import torch

from conftest import load_golden
from mmqg.dims import Dims
from mmqg.synth import make_params, make_batch
from oracle import mmqg_oracle as O
from oracle import ref_loop as R


def rel(a, b):
    return O.rel_err(a, b)


@pytest.mark.parametrize("name", ["small_a", "small_b"])
def test_oracle_matches_reference_small(name):
    fx = load_golden(name)
    d = Dims(**fx["dims"])
    loss, grads = O.loss_and_grads(fx["params"], fx["batch"], d.L, d.TM, d.AM, torch.float64)
    assert abs(float(loss) - float(fx["loss"])) < 1e-10 * abs(float(fx["loss"]))
    assert set(grads) == set(fx["grads"])
    for k, g in fx["grads"].items():
        assert O.rel_err(grads[k], g) < 1e-9, k


@pytest.mark.parametrize("name", ["small_a", "small_b"])
def test_oracle_steps_match_reference(name):
    fx = load_golden(name)
    d = Dims(**fx["dims"])
    p = {k: v.double() for k, v in fx["params"].items()}
    b = {k: (v.double() if v.is_floating_point() else v) for k, v in fx["batch"].items()}
    with torch.no_grad():
        _, aux = O.teacher_forced_loss(p, b, d.L, d.TM, d.AM, return_steps=True)
    for t, st in enumerate(aux["steps"]):
        assert torch.allclose(st["logits"], fx["step_logits"][:, t], rtol=1e-9, atol=1e-11)
        attn = torch.cat([st["a_txt"], st["a_aud"], st["a_vid"]], 1)
        assert torch.allclose(attn, fx["step_attn"][:, t], rtol=1e-9, atol=1e-12)
    # Q1: the length mask is a no-op -- padded slots carry probability mass
    assert float(aux["steps"][0]["a_txt"][:, d.T_t:].sum()) > 0


@pytest.mark.parametrize("name", ["small_a", "small_b", "full_dim"])
def test_oracle_greedy_matches_reference(name):
    fx = load_golden(name)
    d = Dims(**fx["dims"])
    gp = make_params(d, seed=fx["seed"], bias_scale=0.1, out_weight_scale=10.0)
    batch = make_batch(d, seed=fx["seed"] + 1000)
    toks, margins = O.greedy_decode(gp, batch, d.L, d.TM, d.AM, fx["greedy_max_len"], return_margins=True)
    assert torch.equal(toks, fx["greedy_tokens"])
    assert torch.allclose(margins, fx["greedy_margins"].double(), rtol=1e-6, atol=1e-9)


def test_oracle_matches_reference_full_dim_fingerprint():
    fx = load_golden("full_dim")
    d = Dims(**fx["dims"])
    params = make_params(d, seed=fx["seed"])
    batch = make_batch(d, seed=fx["seed"] + 1000)
    loss, grads = O.loss_and_grads(params, batch, d.L, d.TM, d.AM, torch.float64)
    assert abs(float(loss) - float(fx["loss"])) < 1e-10 * abs(float(fx["loss"]))
    for k, fp in fx["grad_fingerprint"].items():
        g = grads[k]
        assert abs(float(g.norm()) - float(fp["norm"])) <= 1e-9 * float(fp["norm"]) + 1e-300, k
        assert torch.allclose(g.flatten()[fp["idx"]], fp["vals"], rtol=1e-8, atol=1e-12), k


def test_ref_loop_port_matches_oracle():
    """The per-sample port (stock torch.nn modules) and the batched restatement agree."""
    fx = load_golden("small_a")
    d = Dims(**fx["dims"])
    loss, grads = R.batch_loss_and_grads(fx["params"], fx["batch"], d.L, torch.float64)
    assert abs(loss - float(fx["loss"])) < 1e-10 * abs(float(fx["loss"]))
    for k, g in fx["grads"].items():
        assert O.rel_err(grads[k], g) < 1e-9, k
    ref = R.RefModules(make_params(d, seed=fx["seed"], bias_scale=0.1, out_weight_scale=10.0), d.L, 0.0, torch.float64)
    b = fx["batch"]
    toks = ref.sample_greedy(b["context"][0], b["frames"][0].double(), b["audio"][0].double(), fx["greedy_max_len"])
    assert toks == fx["greedy_tokens"][0].tolist()


def test_oracle_fp32_noise_floor():
    """fp32 oracle vs fp64 oracle: the self-noise the 1e-3 bar sits on."""
    fx = load_golden("small_a")
    d = Dims(**fx["dims"])
    l32, g32 = O.loss_and_grads(fx["params"], fx["batch"], d.L, d.TM, d.AM, torch.float32)
    assert abs(float(l32) - float(fx["loss"])) < 1e-5 * abs(float(fx["loss"]))
    for k, g in fx["grads"].items():
        assert O.rel_err(g32[k], g) < 1e-4, k


def test_oracle_dropout_masks_semantics():
    """drop=ones is the eval path; the masks scale the input of layers 1.. (torch.nn.LSTM
    inter-layer dropout), so the result equals nn.LSTM fed the same masked activations."""
    import torch
    from mmqg.dims import Dims
    from mmqg.synth import make_batch, make_params
    from oracle import mmqg_oracle as O
    d = Dims(B=3, T_t=4, T_v=2, T_q=3, V=23, E=6, H=8, L=3, H_a=4, H_v=8, F_v=5, TM=6, AM=3)
    params, batch = make_params(d, seed=3), make_batch(d, seed=4)
    ones = {"text": torch.ones(d.L - 1, d.T_t, d.B, d.H), "dec": torch.ones(d.L - 1, d.T_q, d.B, d.H)}
    l0, g0 = O.loss_and_grads(params, batch, d.L, d.TM, d.AM)
    l1, g1 = O.loss_and_grads(params, batch, d.L, d.TM, d.AM, drop=ones)
    assert float(l0) == float(l1)
    assert all(torch.equal(g0[k], g1[k]) for k in g0)
    gen = torch.Generator().manual_seed(0)
    drop = {k: (torch.rand(v.shape, generator=gen) >= 0.2).float() / 0.8 for k, v in ones.items()}
    l2, g2 = O.loss_and_grads(params, batch, d.L, d.TM, d.AM, drop=drop)
    assert float(l2) != float(l0)
    # cross-check the text stack against a manual layer-by-layer nn.LSTM evaluation with the same masks
    p64 = {k: v.double() for k, v in params.items()}
    x = p64["emb.weight"][batch["context"]].transpose(0, 1)          # (T,B,E)
    for l in range(d.L):
        lstm = torch.nn.LSTM(x.shape[2], d.H, 1).double()
        with torch.no_grad():
            lstm.weight_ih_l0.copy_(p64[f"text.lstm.weight_ih_l{l}"]); lstm.weight_hh_l0.copy_(p64[f"text.lstm.weight_hh_l{l}"])
            lstm.bias_ih_l0.copy_(p64[f"text.lstm.bias_ih_l{l}"]); lstm.bias_hh_l0.copy_(p64[f"text.lstm.bias_hh_l{l}"])
            x, _ = lstm(x)
            if l < d.L - 1:
                x = x * drop["text"][l].double()
    masks = [[drop["text"][l, t].double() for l in range(d.L - 1)] for t in range(d.T_t)]
    mem, _, _ = O.text_encode(p64, batch["context"], d.L, d.TM, masks)
    assert torch.allclose(mem[:, :d.T_t].transpose(0, 1), x, atol=1e-12)


def test_oracle_variable_lengths_equal_the_per_sample_loop_of_stock_modules():
    """teacher_forced_loss_varlen (what the batched CUDA path is held to) against the per-sample
    loop over stock torch.nn modules (oracle/ref_loop.py, the train.py:153-177 call structure) with
    every sample cut to its own lengths: loss and every gradient."""
    import torch
    from mmqg.dims import Dims
    from mmqg.synth import make_batch, make_params
    from oracle import mmqg_oracle as O
    from oracle.ref_loop import RefModules
    d = Dims(B=4, T_t=6, T_v=3, T_q=4, V=31, E=6, H=8, L=2, H_a=4, H_v=8, F_v=5, TM=7, AM=4)
    params, batch = make_params(d, seed=13), make_batch(d, seed=14)
    batch["ctx_len"] = torch.tensor([6, 1, 3, 4])
    batch["tgt_len"] = torch.tensor([4, 2, 1, 3])
    batch["n_frames"] = torch.tensor([3, 1, 2, 2])
    loss, grads = O.loss_and_grads(params, batch, d.L, d.TM, d.AM, torch.float64)
    ref = RefModules(params, d.L, 0.0, torch.float64)
    ref.zero_grad()
    total = 0.0
    for b in range(d.B):
        cl, tl, nf = int(batch["ctx_len"][b]), int(batch["tgt_len"][b]), int(batch["n_frames"][b])
        l = ref.sample_loss(batch["context"][b, :cl], batch["frames"][b, :nf].double(), batch["audio"][b, :nf].double(),
                            batch["target"][b, :tl]) / d.B
        l.backward()
        total += float(l.detach())
    assert abs(float(loss) - total) < 1e-10 * abs(total)
    for n, p in ref.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert O.rel_err(grads[n], g) < 1e-9, n


def test_oracle_variable_lengths_match_the_reference_fixture():
    """tests/golden/varlen_a.pt was produced by the reference's own modules driven per sample with
    every sample cut to its own lengths (make_golden.py run_varlen_case): the oracle's
    variable-length path must reproduce its per-sample losses and every gradient."""
    import torch
    from conftest import load_golden
    from mmqg.dims import Dims
    from oracle import mmqg_oracle as O
    fx = load_golden("varlen_a")
    d = Dims(**fx["dims"])
    batch = dict(fx["batch"])
    batch["ctx_len"], batch["tgt_len"], batch["n_frames"] = fx["ctx_len"], fx["tgt_len"], fx["n_frames"]
    loss, grads = O.loss_and_grads(fx["params"], batch, d.L, d.TM, d.AM, torch.float64)
    assert abs(float(loss) - float(fx["loss"])) < 1e-9 * abs(float(fx["loss"]))
    for k, g in fx["grads"].items():
        assert O.rel_err(grads[k], g) < 1e-9, k
    # and the lengths matter: the full-length loss of the same batch is different
    full, _ = O.loss_and_grads(fx["params"], fx["batch"], d.L, d.TM, d.AM, torch.float64)
    assert abs(float(full) - float(fx["loss"])) > 0.1


def test_oracle_adam_wiring_matches_the_reference_fixture():
    """tests/golden/adam_a.pt: three iterations of the reference's own loop with its three Adam
    optimisers (train.py:265-267, 149-181).  The oracle + torch.optim.Adam wired as tests/test_gpu_adam.py
    wires it (the CUDA path's yardstick) must land on the same parameters -- in particular the
    shared embedding, which two of the optimisers step."""
    import torch
    from conftest import load_golden
    from mmqg.dims import Dims
    from oracle import mmqg_oracle as O
    fx = load_golden("adam_a")
    d = Dims(**fx["dims"])
    p = {k: v.detach().double().clone().requires_grad_(True) for k, v in fx["params"].items()}
    emb = p["emb.weight"]
    groups = ([v for k, v in p.items() if k.startswith("video.")], [emb] + [v for k, v in p.items() if k.startswith("text.")],
              [emb] + [v for k, v in p.items() if k.startswith("dec.")])
    opts = [torch.optim.Adam(g, lr=1e-4) for g in groups]
    for it, b in enumerate(fx["batches"]):
        bb = {k: (v.double() if v.is_floating_point() else v) for k, v in b.items()}
        for o in opts:
            o.zero_grad()
        loss = O.teacher_forced_loss(p, bb, d.L, d.TM, d.AM)
        loss.backward()
        for o in opts:
            o.step()
        assert abs(float(loss) - float(fx["losses"][it])) < 1e-9 * abs(float(loss))
    for k, v in fx["final_params"].items():
        upd_ref = v.double() - fx["params"][k].double()
        upd = p[k].detach() - fx["params"][k].double()
        assert float((upd - upd_ref).norm()) <= 1e-7 * float(upd_ref.norm()) + 1e-15, k


@pytest.mark.parametrize("name", ["convstack_a", "convstack_b"])
def test_convstack_oracle_matches_reference_module(name):
    """f2: oracle/convstack_oracle.py against the reference's own VideoConvLstmEncoder (train-mode outputs, every gradient,
    BatchNorm running buffers, eval-mode outputs) at 1e-9."""
    from oracle import convstack_oracle as CO
    fx = load_golden(name)
    out, loss, grads, buffers = CO.loss_and_grads(fx)
    assert abs(loss - fx["loss"]) < 1e-9 * max(1.0, abs(fx["loss"]))
    assert rel(out, fx["out_train"].double()) < 1e-6          # fixtures are stored in fp32
    for k, g in fx["grads"].items():
        assert rel(grads[k], g.double()) < 1e-6, k
    for k, v in buffers.items():
        assert rel(v, fx["bn_after"][k].double()) < 1e-6, k
    p = {k: v.double() for k, v in fx["state0"].items()}
    p.update({k: v.double() for k, v in fx["bn_after"].items() if "running" in k})
    out_eval, _ = CO.video_conv_lstm(fx["frames"].double(), p, False)
    assert rel(out_eval, fx["out_eval"].double()) < 1e-6
