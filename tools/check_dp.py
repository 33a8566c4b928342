"""Data-parallel correctness on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_dp.py [--mode fp32|bf16]

Every rank runs TrainEngine.step on its shard with grad_scale = B_local/B_global and the
bucketed NCCL all-reduce (mmqg.dp.GradReducer); rank 0 additionally runs the whole global
batch on one GPU.  The reduced gradients must match the single-GPU gradients."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-modal-qg_b200"))

from mmqg.dims import Dims  # noqa: E402
from mmqg.dp import GradReducer, shard_batch  # noqa: E402
from mmqg.engine import TrainEngine  # noqa: E402
from mmqg.synth import make_batch, make_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="fp32")
    ap.add_argument("--grad-comm", default="fp32", choices=["fp32", "bf16"], help="element type of the gradient all-reduce")
    ap.add_argument("--reduce", default="overlap", choices=["overlap", "late", "multimem"], help="GradReducer schedule")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bl = 128
    dl = Dims(B=Bl, T_t=20, T_v=4, T_q=6, V=3000, E=300, H=512, L=3, H_a=128, H_v=512, F_v=2048, TM=283, AM=101)
    dg = Dims(**{**dl.asdict(), "B": Bl * world})
    params = make_params(dl, seed=0)
    gbatch = make_batch(dg, seed=77)
    eng = TrainEngine(dl, params, device=dev, mode=args.mode)
    red = GradReducer(eng, world, comm_dtype=torch.bfloat16 if args.grad_comm == "bf16" else torch.float32, schedule=args.reduce)
    local_batch = eng.to_device(shard_batch(gbatch, rank, world))
    graph = None
    # the bf16 exchange lives on the event-driven path only (after_backward); on_phase() always sums in fp32
    for overlap in (("events", "graph") if args.grad_comm == "bf16" or args.reduce != "overlap" else ("events", "graph", True, False)):
        if overlap == "events":                     # event-driven overlap (mmqg_train_backward_events)
            loss = eng.step_dp(local_batch, red, 1.0 / world)
            red.finish()
        elif overlap == "graph":                    # the same step captured, NCCL all-reduces included
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eng.step_dp(local_batch, red, 1.0 / world)
                red.finish()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                eng.step_dp(local_batch, red, 1.0 / world)
                red.finish()
            for g in eng.grad_buckets:
                g.zero_()
            graph.replay()
            graph.replay()
            loss = eng.loss
        elif overlap:
            loss = eng.step(local_batch, grad_scale=1.0 / world, on_phase=red.on_phase)
            red.finish()
        else:                                       # same step, all-reduce only after the whole backward
            loss = eng.step(local_batch, grad_scale=1.0 / world)
            for i in range(4):
                red.on_phase(i)
            red.finish()
        torch.cuda.synchronize()
        gl = loss.clone() / world
        dist.all_reduce(gl)
        grads = {k: v.clone() for k, v in eng.grads.items()}
        if rank == 0:
            ref = TrainEngine(dg, params, device=dev, mode=args.mode)
            ref_loss = float(ref.step(ref.to_device(gbatch)))
            torch.cuda.synchronize()
            worst = max((float((grads[k] - g).norm() / g.norm().clamp_min(1e-30)), k) for k, g in ref.grads.items())
            tol = 1e-3 if args.mode == "fp32" else 3e-2
            if args.grad_comm == "bf16":
                tol = max(tol, 1e-2)            # every shard's gradient is rounded to bf16 and summed in bf16
            print(f"dp world={world} mode={args.mode} comm={args.grad_comm} overlap={overlap}: loss {float(gl):.6f} vs single-GPU {ref_loss:.6f}; "
                  f"worst grad rel err {worst[0]:.2e} ({worst[1]})", flush=True)
            assert abs(float(gl) - ref_loss) < tol * abs(ref_loss)
            assert worst[0] < tol, worst
            del ref
    print(f"rank {rank}: all variants done", flush=True)
    import threading
    threading.Timer(20.0, lambda: os._exit(0)).start()      # a stuck communicator must not hang a finished check
    torch.cuda.synchronize()
    del graph
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
