"""Data-parallel host logic on CPU: world_size-2 gloo processes.

The batch is sharded by sample, every rank scales its gradients by B_local/B_global (what
TrainEngine.step(grad_scale=...) does inside the kernels) and GradReducer sums the buckets in
readiness order.  With the CPU oracle standing in for the kernels, the reduced gradients must
equal the oracle's gradients on the global batch (DP tier)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEngine:
    """grad_buckets / grads laid out exactly like mmqg.engine.TrainEngine, on the CPU."""

    def __init__(self, shapes, L):
        from mmqg.engine import grad_group
        self.grads, self.grad_buckets = {}, []
        offs, total, ranges = {}, 0, []
        for g in range(4 + L):
            lo = total
            for n in [n for n in shapes if grad_group(n, L) == g]:
                offs[n] = total
                total += (int(torch.tensor(shapes[n]).prod()) + 63) // 64 * 64
            ranges.append((lo, total))
        self.flat_grads = torch.zeros(total, dtype=torch.float32)      # one flat buffer, buckets = consecutive ranges of it
        for n, o in offs.items():
            numel = int(torch.tensor(shapes[n]).prod())
            self.grads[n] = self.flat_grads[o:o + numel].view(shapes[n])
        self.grad_buckets = [self.flat_grads[lo:hi] for lo, hi in ranges]


def _worker(rank, world, port, ret, schedule="overlap"):
    for p in (ROOT, os.path.join(ROOT, "multi-modal-qg_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmqg.dims import Dims, param_shapes
    from mmqg.dp import GradReducer, shard_batch
    from mmqg.synth import make_batch, make_params
    from oracle import mmqg_oracle as O
    torch.set_num_threads(2)
    dg = Dims(B=4, T_t=5, T_v=2, T_q=3, V=23, E=8, H=16, L=2, H_a=4, H_v=16, F_v=6, TM=7, AM=3)
    params = make_params(dg, seed=3)
    gbatch = make_batch(dg, seed=4)
    local = shard_batch(gbatch, rank, world)
    assert local["context"].shape[0] == dg.B // world
    loss_l, grads_l = O.loss_and_grads(params, local, dg.L, dg.TM, dg.AM, torch.float64)
    eng = _FakeEngine(param_shapes(dg), dg.L)
    scale = (dg.B // world) / dg.B
    for k, g in grads_l.items():
        eng.grads[k].copy_((scale * g).float())
    red = GradReducer(eng, world, schedule=schedule)
    if schedule == "late":                       # one all-reduce of the flat buffer after the backward
        red.after_backward()
    else:
        for phase in range(4):                   # backward phases: loss head, decoder, video, text layers + embedding
            red.on_phase(phase)
    red.finish()
    loss_g, grads_g = O.loss_and_grads(params, gbatch, dg.L, dg.TM, dg.AM, torch.float64)
    worst = max(O.rel_err(eng.grads[k], g) for k, g in grads_g.items())
    lsum = torch.tensor([float(loss_l) * scale], dtype=torch.float64)
    dist.all_reduce(lsum)
    if rank == 0:
        ret["worst"] = worst
        ret["loss_err"] = abs(float(lsum) - float(loss_g)) / abs(float(loss_g))
    dist.destroy_process_group()


@pytest.mark.parametrize("schedule", ["overlap", "late"])
def test_dp_two_ranks_match_global_batch(schedule):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000 + (7 if schedule == "late" else 0)
    mp.spawn(_worker, args=(world, port, ret, schedule), nprocs=world, join=True)
    assert ret["worst"] < 1e-5, dict(ret)
    assert ret["loss_err"] < 1e-12


def test_grad_groups_cover_every_parameter_once():
    from mmqg.dims import Dims, param_shapes
    from mmqg.engine import grad_group
    d = Dims(B=1, T_t=2, T_v=1, T_q=2, V=11, E=4, H=8, L=3, H_a=4, H_v=8, F_v=4, TM=3, AM=2)
    groups = {n: grad_group(n, d.L) for n in param_shapes(d)}
    assert set(groups.values()) == set(range(4 + d.L))
    assert groups["dec.out_layer.weight"] == 0 and groups["dec.lstm.weight_hh_l2"] == 1
    assert groups["video.lstm.bias_ih_l0"] == 2
    # text layers in BPTT completion order (top layer first), the shared embedding last
    assert groups["text.lstm.weight_ih_l2"] == 3 and groups["text.lstm.bias_hh_l1"] == 4 and groups["text.lstm.weight_ih_l0"] == 5
    assert groups["emb.weight"] == 6
