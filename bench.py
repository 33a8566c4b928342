#!/usr/bin/env python
"""Benchmark of the multi-modal-qg training hot path on B200 (contract: task statement (4)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]

A "step" is one teacher-forced forward + backward of the encoder/decoder over one synthetic
batch (BASELINE.json configs[1]: B=256 per GPU, T_t=100, T_v=10x2048, T_a=10x128, T_q=20,
V=10k).  Rank 0 prints ONE JSON line.  Multi-GPU runs are launched with torch.distributed.run
(one process per GPU, NCCL); the batch is sharded by sample, gradients are all-reduced in the
order they become final, overlapped with the rest of the backward pass.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "multi-modal-qg_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch override")
    ap.add_argument("--mode", default=os.environ.get("MMQG_MODE", "bf16"), choices=["fp32", "bf16", "fp32_tc"],
                    help="bf16 = tcgen05 path (BASELINE configs[1] names bf16); fp32 = SIMT parity mode; fp32_tc = the parity "
                         "path with every contraction on tcgen05 (operands split into bf16 hi/lo, three products)")
    ap.add_argument("--cpu-samples", type=int, default=32, help="samples the CPU baseline leg times")
    ap.add_argument("--dropout", type=float, default=None,
                    help="inter-layer LSTM dropout (config.py text_lstm_dropout = dec_lstm_dropout); default 0.2 in "
                         "bf16 mode (the reference's train-mode value), 0 in the fp32 parity mode")
    ap.add_argument("--adam", action="store_true",
                    help="also run the fused Adam step (SURVEY section 8 f1) inside every step; off by default: the "
                         "metric is forward + backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--probe", type=int, default=None, help="kernel class the roofline probe times (see mmqg.h)")
    ap.add_argument("--greedy", action="store_true",
                    help="time the greedy decode path (BASELINE.json configs[4]: B=1024, 30 tokens) instead of the train "
                         "step; implied by --config 5")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check of the timed configuration")
    ap.add_argument("--grad-comm", default="fp32", choices=["fp32", "bf16", "auto"],
                    help="N > 1: element type of the gradient all-reduce (fp32 = exact sums; bf16 = buckets rounded to bf16 for the "
                         "exchange, half the NVLink bytes; auto = bf16 in bf16 mode)")
    ap.add_argument("--reduce", default=os.environ.get("MMQG_REDUCE", "multimem"), choices=["overlap", "late", "multimem"],
                    help="N > 1: overlap = one NCCL all-reduce per gradient group under the rest of the backward; late = one all-reduce "
                         "of the flat gradient buffer after the backward; multimem = per group, our own reduce kernel over NVSwitch "
                         "multicast memory (mmqg_allreduce_multimem; fp32 sums; falls back to overlap on every rank if any rank has no multicast mapping)")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def flops_per_sample(d):
    """Algorithmic FLOPs of one forward; train = 3x."""
    G = 4 * d.H
    text = d.T_t * 2 * G * ((d.E + d.H) + (d.L - 1) * 2 * d.H)
    video = d.T_v * 2 * 4 * d.H_v * (d.F_v + d.H_v)
    dec = d.T_q * (2 * (d.E + d.H) * (d.TM + 2 * d.AM) + 2 * (d.T_t * d.H + d.T_v * d.H_a + d.T_v * d.H_v)
                   + 2 * G * ((d.E + d.H + d.H_a + d.H_v) + d.H + (d.L - 1) * 2 * d.H) + 2 * d.H * d.V)
    return text + video + dec


def cpu_reference_samples_per_s(d, n_samples, seed=0, dropout_p=0.0):
    """The reference's per-sample train iteration (train.py:149-177: zero_grad, encoder,
    teacher-forced decoder, loss.backward()) on the host cores, via the oracle's per-sample
    port (stock torch.nn modules, same call granularity).  The reference itself cannot
    travel to the GPU box (/root/reference is absent there), hence kind = "port"."""
    from mmqg.dims import Dims
    from mmqg.synth import make_batch, make_params
    from oracle.ref_loop import RefModules, train_samples
    dd = Dims(**{**d.asdict(), "B": n_samples})
    params = make_params(dd, seed=seed)
    batch = make_batch(dd, seed=1234)
    ref = RefModules(params, dd.L, dropout_p=dropout_p)
    ref.train()
    train_samples(ref, batch, 1)                       # warm-up (oneDNN primitive caches)
    t0 = time.perf_counter()
    train_samples(ref, batch, n_samples)
    dt = time.perf_counter() - t0
    return n_samples / dt, dt


def run_reference(args, d, rank, world):
    """--impl reference: the CPU path timed on the host cores, same metric/config keys."""
    if rank != 0:
        return
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the reference arm is the CPU
    # implementation "with all the host threads it can use", so take the cores back explicitly
    # (the other ranks have already returned and do no work).
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, ncpu))
    # bounded sample: about 200 samples in total (~100 s at ~2 samples/s on 16 cores), whatever K is
    per_step = max(1, min(8, args.cpu_samples, 200 // max(1, args.steps)))
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_samples_per_s(d, 1, dropout_p=args.dropout)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        sps, dt = cpu_reference_samples_per_s(d, per_step, dropout_p=args.dropout)
        t_total += dt
        n_total += per_step
    v = n_total / t_total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "train samples/sec (fwd+bwd)", "value": v, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(d, args, world),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
                         "sample": f"{per_step} samples per step x {args.steps} steps of the per-sample reference loop "
                                   f"(train.py:149-177 semantics, stock torch.nn modules) at the config's lengths; "
                                   f"per-sample cost is independent of B"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def committed_traffic(kernel_class, launches_per_step):
    """DRAM bytes per launch of the probed kernel class from the newest committed ncu --set full capture
    (profiles/rNN_traffic.json; dram__bytes_read.sum + dram__bytes_write.sum of the FINAL kernels, summed
    over one step), or None when no capture covers that class.  Not measurable live: ncu replays."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f)
            if t.get("kernel_class") == kernel_class and launches_per_step > 0:
                return {"bytes_per_launch": t["bytes_per_step"] / launches_per_step, "source": os.path.basename(path),
                        "algorithmic_bytes_per_launch": t.get("algorithmic_bytes_per_step", 0) / launches_per_step}
        except (OSError, ValueError, KeyError):
            continue
    return None


def workload_config(d, args, world):
    what = (f"greedy decode {d.T_q} tokens" if args.greedy else "teacher-forced train step fwd+bwd")
    return {"workload": f"BASELINE.json configs[{args.config - 1}]: {what}, "
                        f"B={d.B}/GPU T_t={d.T_t} T_v={d.T_v}x{d.F_v} T_a={d.T_v}x{d.H_a} T_q={d.T_q} V={d.V} "
                        f"E={d.E} H={d.H} L={d.L} TM={d.TM} AM={d.AM}",
            "global_batch": d.B * world, "per_gpu_batch": d.B, "parallelism": f"dp{world}",
            "dropout_p": 0.0 if args.greedy else args.dropout,
            "optimizer": "fused adam" if args.adam else "none (metric is fwd+bwd)",
            "mode": args.mode, "cuda_graph": not args.no_graph,
            **({"grad_comm": ("bf16" if args.grad_comm == "bf16" or (args.grad_comm == "auto" and args.mode == "bf16") else "fp32"),
                "grad_reduce": args.reduce} if world > 1 else {}),
            "l2": "working set (2 GB activations + 111 MB weights per step) exceeds the 126 MB L2; no flush needed"}


DTYPE_OF_MODE = {"fp32": "f32", "bf16": "bf16", "fp32_tc": "bf16x3 (f32 operands split hi/lo, f32 accumulation)"}


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def parity_check(eng, d, params, host_batch, mode, config_n):
    """One dropout-free step of the TIMED configuration against the CPU oracle (test infrastructure used as
    the checker, like the cpu_baseline leg): bf16 mode vs the oracle with identically bf16-rounded weights
    (bar: loss 5e-3, gradients 5e-2 per tensor), fp32 mode vs the plain oracle (bar 1e-3)."""
    from mmqg.synth import round_params_bf16
    from oracle import mmqg_oracle as O
    odt = torch.float32 if config_n == 4 else torch.float64     # cfg-4 in fp64 takes minutes; fp32 oracle noise is 1e-6
    t0 = time.perf_counter()
    p = round_params_bf16(params) if mode == "bf16" else params
    loss_ref, grads_ref = O.loss_and_grads(p, host_batch, d.L, d.TM, d.AM, odt)
    keep = eng.dropout_p
    eng.dropout_p = 0.0
    try:
        loss = float(eng.step(eng.to_device(host_batch)))
        torch.cuda.synchronize()
    finally:
        eng.dropout_p = keep
    errs = {k: rel_err(eng.grads[k], g) for k, g in grads_ref.items()}
    worst = max(errs.items(), key=lambda kv: kv[1])
    lt, gt = (5e-3, 5e-2) if mode == "bf16" else (1e-3, 1e-3)
    loss_rel = abs(loss - float(loss_ref)) / abs(float(loss_ref))
    return {"loss": loss, "oracle_loss": float(loss_ref), "loss_rel": loss_rel, "worst_grad_rel": worst[1], "tensor": worst[0],
            "median_grad_rel": sorted(errs.values())[len(errs) // 2], "loss_tol": lt, "grad_tol": gt,
            "ok": bool(loss_rel < lt and worst[1] < gt), "dropout_p": 0.0,
            "oracle": f"oracle/mmqg_oracle.py {str(odt).split('.')[-1]}" + (", bf16-rounded weights" if mode == "bf16" else ""),
            "oracle_seconds": time.perf_counter() - t0}


def rank_seed(rank):
    """Base dropout seed of a rank: every rank draws its own masks (a sample's mask must not repeat across
    the global batch, as it would if all ranks used seed 0 with the lock-step call counter)."""
    return (rank * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF


def dp_check(eng, d, params, reducer, world, rank, dev, dbatch):
    """Data-parallel correctness, visible to the driver: one sharded step (dropout as timed, every rank with
    its own mask seed) with the bucketed NCCL all-reduce on every rank; rank 0 then runs EVERY rank's shard
    locally (same kernels, that rank's seed and call counter, grad_scale 1/world) and sums the gradients --
    the global batch on one GPU, shard by shard.  Returns the worst relative error of the reduced gradient
    tensors (fp32 sums in a different order: expected ~1e-6)."""
    import torch.distributed as dist
    from mmqg.synth import make_batch
    CALLS = 7777                                   # common value of the device-side dropout call counter
    eng.reset_dropout_calls(CALLS)
    loss = eng.step_dp(dbatch, reducer, 1.0 / world)
    reducer.finish()
    torch.cuda.synchronize()
    reduced = eng.flat_grads.clone()
    gl = loss.clone() / world
    dist.all_reduce(gl)
    out = None
    if rank == 0:
        acc = torch.zeros_like(reduced)
        lsum = 0.0
        for r in range(world):
            b = eng.to_device(make_batch(d, seed=1234 + r))
            eng.seed = rank_seed(r)
            eng.reset_dropout_calls(CALLS)
            lsum += float(eng.step(b, grad_scale=1.0 / world)) / world
            torch.cuda.synchronize()
            acc += eng.flat_grads
        eng.seed = rank_seed(0)
        # fp32 exchange: the same numbers summed in another order; bf16 exchange: every shard's gradient rounded to bf16
        # (2^-9 relative per element) and summed in bf16 by NCCL
        tol = 1e-3 if reducer.stage is None else 2e-2
        worst = (0.0, None)
        for k, g in eng.grads.items():
            o, n = eng.offsets[k], g.numel()
            e = rel_err(reduced[o:o + n], acc[o:o + n])
            if e > worst[0]:
                worst = (e, k)
        out = {"worst_grad_rel": worst[0], "tensor": worst[1], "loss_global": float(gl), "loss_local_sum": lsum,
               "loss_rel": abs(float(gl) - lsum) / abs(lsum), "ok": bool(worst[0] < tol), "tol": tol,
               "grad_comm": "bf16" if reducer.stage is not None else "fp32", "dropout_p": eng.dropout_p,
               "what": f"all-reduced gradients of {world} ranks vs the same {world} shards (each with its rank's dropout "
                       f"seed) run one after another on rank 0"}
    dist.barrier()
    return out


def persistent_union(eng, dbatch):
    """Wall-clock share of the persistent recurrent kernels inside one eager step: union of their
    [start, end] intervals (%globaltimer stamps written by the kernels, debug hook mmqg_debug_ktrace)
    over the step's event-timed duration, and their mean concurrency (summed time / union)."""
    from mmqg import _cabi
    L = C.CDLL(_cabi.LIB_PATH)
    L.mmqg_debug_ktrace.argtypes = [C.c_void_p]
    buf = torch.zeros(1 + 3 * 1024, dtype=torch.int64, device=eng.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.mmqg_debug_ktrace(buf.data_ptr())
    try:
        e0.record()
        eng.step(dbatch)
        e1.record()
        torch.cuda.synchronize()
    finally:
        L.mmqg_debug_ktrace(None)
    t = buf.cpu()
    n = int(t[0])
    iv = sorted((int(t[2 + 3 * i]), int(t[3 + 3 * i])) for i in range(n))
    if not iv:
        return None
    union, cur_s, cur_e = 0, iv[0][0], iv[0][1]
    for s_, e_ in iv[1:]:
        if s_ > cur_e:
            union += cur_e - cur_s
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    union += cur_e - cur_s
    summed = sum(e_ - s_ for s_, e_ in iv)
    step_ms = e0.elapsed_time(e1)
    return {"union_ms": union / 1e6, "summed_ms": summed / 1e6, "eager_step_ms": step_ms, "launches": n,
            "share_of_step": (union / 1e6) / step_ms, "mean_concurrency": summed / union}


def run_greedy(args, d, rank, world, local, dev):
    """BASELINE.json configs[4]: greedy decode, B=1024, 30 tokens (evaluate.py:45-104 / train.py:81-110).
    value = samples/s with the batch resident; e2e = host batch in, host tokens out every step."""
    from mmqg.engine import TrainEngine, launch_count
    from mmqg.synth import make_batch, make_params, round_params_bf16
    params = make_params(d, seed=0, bias_scale=0.1, out_weight_scale=10.0)      # input-sensitive weights
    eng = TrainEngine(d, params, device=dev, mode=args.mode)
    host = make_batch(d, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items() if k != "target"}
    db = eng.to_device({k: v for k, v in host.items() if k != "target"})
    n0 = launch_count()
    toks = eng.greedy(db, d.T_q)
    torch.cuda.synchronize()
    launches_per_step = launch_count() - n0
    g = None
    if not args.no_graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            toks = eng.greedy(db, d.T_q)

    def run():
        """One decode of the resident batch; returns the (B, T_q) token tensor on the device."""
        if g is not None:
            g.replay()
            return toks
        return eng.greedy(db, d.T_q)

    for _ in range(max(args.warmup, 3)):
        run()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # end to end: pinned host batch -> device, decode, tokens -> pinned host memory, every step
    host_toks = torch.empty(d.B, d.T_q, dtype=torch.int64).pin_memory()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        for k, v in pinned.items():
            db[k].copy_(v, non_blocking=True)
        host_toks.copy_(run(), non_blocking=True)
        torch.cuda.synchronize()
    e3.record()
    torch.cuda.synchronize()
    ms_e2e = e2.elapsed_time(e3)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    line = {"metric": "greedy decode samples/sec", "value": d.B * args.steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_OF_MODE[args.mode], "data": "synthetic",
            "config": workload_config(d, args, world), "clocks": clocks,
            "e2e": {"value": d.B * args.steps / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": host_toks.numel() * 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_per_step * args.steps)}
    if not args.no_parity:
        from oracle import mmqg_oracle as O
        p = round_params_bf16(params) if args.mode == "bf16" else params
        want, margins = O.greedy_decode(p, host, d.L, d.TM, d.AM, d.T_q, torch.float64, return_margins=True)
        got = eng.greedy(db, d.T_q).cpu()
        thr = 0.5 if args.mode == "bf16" else 1e-4
        safe = (margins > thr).long().cumprod(1).bool()
        line["parity"] = {"token_match_rate": float((got == want).float().mean()),
                          "rows_exact": int((got == want).all(1).sum()), "rows": d.B,
                          "exact_where_margin_above": thr, "resolvable_positions": float(safe.float().mean()),
                          "ok": bool(torch.equal(got[safe], want[safe])),
                          "oracle_margin_min": float(margins.min()), "oracle_margin_median": float(margins.median()),
                          "oracle": "oracle/mmqg_oracle.py float64" + (", bf16-rounded weights" if args.mode == "bf16" else "")}
    if not args.no_cpu_baseline:
        from oracle.ref_loop import RefModules
        ref = RefModules(params, d.L, dropout_p=0.0)
        ref.train(False)
        n = min(16, d.B)
        ref.sample_greedy(host["context"][0], host["frames"][0], host["audio"][0], d.T_q)
        t0 = time.perf_counter()
        for b in range(n):
            ref.sample_greedy(host["context"][b], host["frames"][b], host["audio"][b], d.T_q)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n} samples through the per-sample greedy loop (train.py:81-110 semantics, stock "
                                          f"torch.nn modules), {dt:.1f} s; host has {os.cpu_count()} cpus"}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.config == 5:
        args.greedy = True
    if args.probe is None:
        args.probe = 7 if args.mode == "bf16" else 1       # dominant kernel class of each mode
    if args.dropout is None:
        args.dropout = 0.2 if args.mode == "bf16" else 0.0
    from mmqg.dims import config as cfg
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    d = cfg(args.config, args.batch)

    if args.impl == "reference":
        run_reference(args, d, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.greedy:
        if rank == 0:
            run_greedy(args, d, rank, 1, local, dev)          # replicas only: the decode path has no exchange step
        return
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from mmqg import _cabi
    from mmqg.engine import TrainEngine, launch_count
    from mmqg.synth import make_batch, make_params
    from mmqg.dp import GradReducer

    params = make_params(d, seed=0)
    eng = TrainEngine(d, params, device=dev, mode=args.mode, dropout_p=args.dropout)
    # every rank draws its own dropout masks (a sample's mask must not repeat across the global batch)
    eng.seed = rank_seed(rank)
    host = make_batch(d, seed=1234 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    dbatch = eng.to_device(host)
    comm_bf16 = args.grad_comm == "bf16" or (args.grad_comm == "auto" and args.mode == "bf16")
    reducer = None
    if world > 1:
        import torch.distributed as dist
        if args.reduce == "multimem" and comm_bf16:
            args.reduce = "overlap"            # the multimem kernel sums in fp32
        if args.reduce == "multimem":          # needs NVSwitch multicast on every rank: agree, else NCCL
            ok = torch.ones(1, device=dev)
            try:
                reducer = GradReducer(eng, world, schedule="multimem")
            except Exception as e:             # noqa: BLE001
                print(f"[bench rank {rank}] multimem all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr, flush=True)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok) == 0.0:
                args.reduce, reducer = "overlap", None
        if reducer is None:
            reducer = GradReducer(eng, world, comm_dtype=torch.bfloat16 if comm_bf16 else torch.float32, schedule=args.reduce)
    gscale = 1.0 / world

    if args.adam:
        eng.adam_init()

    def eager_step(b):
        if reducer:
            loss = eng.step_dp(b, reducer, gscale)     # all-reduce per gradient group behind its ready event
            reducer.finish()
        else:
            loss = eng.step(b)
        if args.adam:
            eng.adam_step(lr=1e-4)
        return loss

    # One step is a few hundred kernel launches; replaying them as a CUDA graph takes the host launch path
    # out of the loop.  With N > 1 the NCCL all-reduces are captured in the same graph.
    use_graph = not args.no_graph
    graph_holder = []
    n_before = launch_count()
    eager_step(dbatch)
    torch.cuda.synchronize()
    launches_per_step = launch_count() - n_before

    def make_step(b):
        """A callable running one step on device batch b (a graph is bound to the buffers it captured)."""
        if not use_graph:
            return lambda: eager_step(b)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager_step(b)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            eager_step(b)
        idx = len(graph_holder)
        graph_holder.append(g)
        del g                                               # teardown() must be able to drop the last reference

        def replay():
            graph_holder[idx].replay()
            return eng.loss
        return replay

    step_resident = make_step(dbatch)

    def one_step(b):
        """b must be dbatch (the graph is bound to its buffers)."""
        return step_resident()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step(dbatch)
    barrier()

    # ---- timed region 1: inputs resident in HBM -------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step(dbatch)
    e1.record()
    barrier()
    launches = launches_per_step * args.steps
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end from pinned host buffers ---------------------------
    # Every step's inputs are copied host -> device from pinned memory and its loss is read back on
    # the host; the copy of step i+1 runs on a copy stream under the compute of step i (HostFeed).
    from mmqg.engine import HostFeed
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    feed = HostFeed(eng, host, make_step)
    for _ in range(2):                                       # warm both buffers / graphs
        feed.prefetch(pinned)
        feed.step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    host_loss = 0.0
    feed.prefetch(pinned)                                    # inputs of the first timed step
    for i in range(args.steps):
        loss = feed.step()
        if i + 1 < args.steps:
            feed.prefetch(pinned)                            # inputs of step i+1, under step i
        host_loss = float(loss.item())                      # device -> host read of the step's result
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    # ---- data-parallel correctness (N > 1), visible in the bench line ------------------------
    print(f"[bench rank {rank}] timed regions done: {ms / args.steps:.3f} ms/step", file=sys.stderr, flush=True)
    dpc = dp_check(eng, d, params, reducer, world, rank, dev, dbatch) if world > 1 else None
    print(f"[bench rank {rank}] dp_check done", file=sys.stderr, flush=True)

    # ---- roofline probe of the dominant kernel class (separate eager pass, same step) ---------
    roof = None
    if rank == 0:
        L = _cabi.lib()
        _cabi.check(L.mmqg_probe_start(args.probe))
        for _ in range(2):
            eng.step(dbatch)                              # local only: no collectives in the probe pass
        torch.cuda.synchronize()
        tms, n, fl, by = C.c_double(), C.c_ulonglong(), C.c_double(), C.c_double()
        _cabi.check(L.mmqg_probe_stop(C.byref(tms), C.byref(n), C.byref(fl), C.byref(by)))
        pk = peaks()
        names = {1: "per-timestep recurrent GEMM (gemm_f32_kernel, B x 4H x H class)", 2: "hoisted whole-sequence GEMM",
                 3: "lstm_pointwise", 4: "attention step", 5: "nll_rows", 6: "embedding",
                 7: "persistent recurrent-cell kernels lstm_seq_fwd/bwd_kernel (tcgen05, one launch per layer, chunk and "
                    "direction; FLOPs = 2*T*B*4H*H per launch)"}
        if n.value and tms.value > 0:
            if args.probe in (1, 2, 7):
                ach = fl.value / (tms.value * 1e-3) / 1e12
                roof = {"bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                        "frac": ach / pk["tflops"]}
            else:
                ach = by.value / (tms.value * 1e-3) / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / pk["hbm_gbs"]}
            tr = committed_traffic(args.probe, n.value / 2)
            roof["traffic"] = tr["bytes_per_launch"] if tr else None
            if tr:
                roof["traffic_source"] = tr["source"]
                roof["traffic_over_algorithmic"] = (tr["bytes_per_launch"] / tr["algorithmic_bytes_per_launch"]
                                                    if tr["algorithmic_bytes_per_launch"] else None)
            roof.update({"kernel": names.get(args.probe), "launches_per_step": n.value // 2,
                         "avg_launch_us": 1e3 * tms.value / n.value,
                         "summed_ms_per_step": tms.value / 2,
                         "peak_source": pk["source"] + (" (sustained bf16 cuBLAS; this kernel is fp32 SIMT)"
                                                        if args.probe in (1, 2) and args.mode == "fp32" else "")})
            if args.probe == 7:
                u = persistent_union(eng, dbatch)
                if u:
                    roof["share_of_step"] = u["share_of_step"]
                    roof["share_note"] = ("wall-clock union of the kernels' [start,end] intervals (%globaltimer) / the same "
                                          "eager step's event time; mean_concurrency = summed kernel time / union")
                    roof["mean_concurrency"] = u["mean_concurrency"]
                    roof["union_ms"] = u["union_ms"]
                    roof["eager_step_ms"] = u["eager_step_ms"]

    if rank != 0:
        teardown(world, graph_holder)
        return
    sps = d.B * world * args.steps / (ms * 1e-3)
    sps_e2e = d.B * world * args.steps / (ms_e2e * 1e-3)
    f_train = 3 * flops_per_sample(d)
    line = {
        "metric": "train samples/sec (fwd+bwd)", "value": sps, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE_OF_MODE[args.mode], "data": "synthetic",
        "config": workload_config(d, args, world),
        "clocks": clocks,
        "e2e": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": host_loss,
                "last_loss_note": "rank 0's local-shard mean with dropout on; parity.loss is the checked, dropout-free value"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "step_tensor_frac": (f_train * sps / world) / 1e12 / peaks()["tflops"],
        "algorithmic_gflop_per_sample": f_train / 1e9,
    }
    if dpc is not None:
        line["dp_check"] = dpc
    if world == 1 and not args.no_parity:
        line["parity"] = parity_check(eng, d, params, host, args.mode, args.config)
    if world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_reference_samples_per_s(d, args.cpu_samples, dropout_p=args.dropout)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_samples} samples of the same workload through the per-sample "
                                          f"reference loop (train.py:149-177 semantics, stock torch.nn modules), "
                                          f"{dt:.1f} s; host has {os.cpu_count()} cpus"}
    print(json.dumps(line), flush=True)
    teardown(world, graph_holder)


def teardown(world, graph_holder):
    """Release the captured graph (it holds NCCL kernels) BEFORE the communicator goes away; the
    result line is already printed, so a communicator that refuses to shut down must not turn a
    finished run into a hang: a timer ends the process with status 0."""
    if world <= 1:
        return
    import threading
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Timer(30.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    torch.cuda.synchronize()
    graph_holder.clear()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
