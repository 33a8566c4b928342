"""CPU oracle for f2, the conv stack of VideoConvLstmEncoder on raw frames.  TEST INFRASTRUCTURE ONLY (see __init__.py).

Restates reference model/encoder.py:58-67 in plain tensor arithmetic, gradients by autograd:
    x = video_frames.view(T, C, H, W)                                   (:59-64, a `view`, not a permute: SURVEY App. B Q4)
    block1 = maxpool1(bn2(relu(conv2(bn1(relu(conv1(x)))))))           (:62)
    block2 = maxpool2(bn4(relu(conv4(bn3(relu(conv3(block1)))))))      (:63)
    lstm(flatten(block2).view(T, 1, -1))                                (:65-67), zero initial state
Convolutions are written as unfold + matmul, BatchNorm2d with explicit batch statistics (train mode: biased variance for
the normalisation, unbiased for the running buffer, momentum 0.1) and MaxPool2d(k, k) as a reshape + max, so nothing here
calls torch's conv / batch_norm / max_pool kernels.  Pinned at 1e-9 to fixtures produced by the reference's own module
(tests/golden/convstack_*.pt, tests/test_oracle_golden.py)."""
import torch

from .mmqg_oracle import lstm_cell


def conv2d(x, w, b, stride=1):
    N, C, H, W = x.shape
    Co, _, K, _ = w.shape
    Ho, Wo = (H - K) // stride + 1, (W - K) // stride + 1
    cols = torch.nn.functional.unfold(x, K, stride=stride)                    # (N, C*K*K, Ho*Wo): pure gather
    y = w.reshape(Co, -1) @ cols + b.view(1, -1, 1)
    return y.view(N, Co, Ho, Wo)


def batchnorm(x, gamma, beta, running_mean, running_var, training, eps=1e-5, momentum=0.1):
    """Returns (y, new_running_mean, new_running_var)."""
    if training:
        mean = x.mean((0, 2, 3))
        var = ((x - mean.view(1, -1, 1, 1)) ** 2).mean((0, 2, 3))
        n = x.shape[0] * x.shape[2] * x.shape[3]
        new_rm = (1 - momentum) * running_mean + momentum * mean.detach()
        new_rv = (1 - momentum) * running_var + momentum * var.detach() * n / (n - 1)
    else:
        mean, var, new_rm, new_rv = running_mean, running_var, running_mean, running_var
    y = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return y, new_rm, new_rv


def maxpool(x, K):
    N, C, H, W = x.shape
    Hp, Wp = (H - K) // K + 1, (W - K) // K + 1
    x = x[:, :, :Hp * K, :Wp * K].reshape(N, C, Hp, K, Wp, K).permute(0, 1, 2, 4, 3, 5).reshape(N, C, Hp, Wp, K * K)
    return x.max(-1).values


def conv_stack(x, p, training, K=3, stride=1):
    """x (N,C,H,W); p: state_dict-like {conv{i}.weight/bias, bn{i}.weight/bias/running_mean/running_var}.
    Returns (features (N,10,h,w), {bn{i}.running_mean/var after the pass})."""
    buffers = {}
    for i in (1, 2, 3, 4):
        x = torch.relu(conv2d(x, p[f"conv{i}.weight"], p[f"conv{i}.bias"], stride))
        x, rm, rv = batchnorm(x, p[f"bn{i}.weight"], p[f"bn{i}.bias"], p[f"bn{i}.running_mean"], p[f"bn{i}.running_var"], training)
        buffers[f"bn{i}.running_mean"], buffers[f"bn{i}.running_var"] = rm, rv
        if i in (2, 4):
            x = maxpool(x, K)
    return x, buffers


def video_conv_lstm(frames, p, training, K=3, stride=1):
    """frames (1,C,T,H,W) as the reference receives them -> (T,1,hidden) LSTM outputs and the BatchNorm buffers."""
    _, C, T, H, W = frames.shape
    feats, buffers = conv_stack(frames.reshape(T, C, H, W), p, training, K, stride)       # encoder.py:64: view, not permute
    feats = feats.reshape(T, -1)
    hidden = p["lstm.weight_hh_l0"].shape[1]
    h = torch.zeros(1, hidden, dtype=frames.dtype)
    c = torch.zeros(1, hidden, dtype=frames.dtype)
    outs = []
    for t in range(T):
        h, c = lstm_cell(feats[t:t + 1], h, c, p["lstm.weight_ih_l0"], p["lstm.weight_hh_l0"], p["lstm.bias_ih_l0"], p["lstm.bias_hh_l0"])
        outs.append(h)
    return torch.stack(outs), buffers


def loss_and_grads(fx, dtype=torch.float64):
    """Train-mode pass over a convstack fixture: (out, loss, {param: grad}, buffers) with loss = sum(out * proj)."""
    p = {k: v.to(dtype).clone().requires_grad_("running" not in k) for k, v in fx["state0"].items() if "num_batches" not in k}
    out, buffers = video_conv_lstm(fx["frames"].to(dtype), p, True)
    loss = (out * fx["proj"].to(dtype)).sum()
    names = [k for k, v in p.items() if isinstance(v, torch.Tensor) and v.requires_grad]
    grads = dict(zip(names, torch.autograd.grad(loss, [p[k] for k in names])))
    return out.detach(), float(loss), grads, {k: v.detach() for k, v in buffers.items()}
