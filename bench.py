#!/usr/bin/env python
"""Benchmark of the TT-LSTM / TT-GRU recurrence path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl ours|reference]

Metric (BASELINE.json): cell-steps/sec = batch x T / seconds of one forward(+backward) pass over
the whole stack.  One "step" = one such pass over one batch of synthetic input of the config's
shape.  Default workload = BASELINE.json configs[1] (permuted-MNIST TT-GRU, T=784, batch 1024,
fwd+bwd).  Batch is per GPU (weak scaling): N GPUs process N x batch sequences; with N > 1 the
TT-core gradients are all-reduced with NCCL inside the step.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events around each
step, L2 flushed between steps, max over ranks); `e2e` = the same pass through the public module
API from pinned HOST input with the H2D copy and a D2H read of the result inside the timed region;
`roofline` = the dominant kernel against the FP32 FFMA peak measured live by an FFMA probe kernel
(MEASURED_PEAKS.json carries no FP32 figure), in ALGORITHMIC FLOPs of the reference's core-by-core
sweep (where the engine contracts the ih projection in the cheaper dense order it executes fewer
FLOPs than credited; `roofline.ih_projection` lists both costs per layer); `cpu_baseline` = the CPU oracle (a restatement of the
reference's PyTorch path; the reference itself cannot travel to the GPU box) timed on the host
cores on a bounded sample.  `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import ctypes as C
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from contextlib import redirect_stdout

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name, cell, I, H, L, d, r, B (per GPU), T, mode, upstream gradient, input kind, seed
CONFIGS = {
    1: dict(name="cfg1 sequential-MNIST TT-LSTM d2 r4", cell="lstm", I=1, H=256, L=1, d=2, r=4, B=256, T=784,
            mode="fwd+bwd", grad="out_last", inp="digits", seed=1111),
    2: dict(name="cfg2 permuted-MNIST TT-GRU d2 r4", cell="gru", I=1, H=256, L=1, d=2, r=4, B=1024, T=784,
            mode="fwd+bwd", grad="out_last", inp="digits_perm", seed=1111),
    3: dict(name="cfg3 GE2E speaker encoder 3xTT-LSTM d3 r8", cell="lstm", I=40, H=256, L=3, d=3, r=8, B=640, T=160,
            mode="fwd+bwd", grad="hT", inp="uniform", seed=11),
    4: dict(name="cfg4 speaker encoder inference 3xTT-LSTM d4 r16", cell="lstm", I=40, H=256, L=3, d=4, r=16,
            B=2048, T=160, mode="fwd", grad=None, inp="uniform", seed=11),
    5: dict(name="cfg5 TT-LSTM H1024 d4 r8 sweep", cell="lstm", I=256, H=1024, L=1, d=4, r=8, B=4096, T=2000,
            mode="fwd+bwd", grad="dense", inp="uniform", seed=11),
}
# bounded CPU sample (batch) per config for the cpu_baseline / reference arm: T stays at the config
# value (the reference's cost is dominated by per-step dispatch and its O(T^2) backward)
CPU_SAMPLE_B = {1: 64, 2: 64, 3: 32, 4: 16, 5: 4}
CPU_SAMPLE_T = {1: 784, 2: 784, 3: 160, 4: 160, 5: 200}


# ------------------------------------------------------------------------------------------
def chain_flops(in_modes, out_modes, ranks):
    """FLOPs per row of the right-to-left TT sweep: sum_k 2 * M_k * K_k * N_k."""
    d = len(in_modes)
    total = 0
    for k in range(d):
        m = int(np.prod(out_modes[k + 1:])) * int(np.prod(in_modes[:k]))
        total += 2 * m * (in_modes[k] * ranks[k + 1]) * (out_modes[k] * ranks[k])
    return total


def algorithmic_flops(cfg):
    """Forward FLOPs per sequence-step for the whole stack, and for the hh chain of one layer."""
    from tensorized_rnn_b200.shapes import tt_shape
    G = 4 if cfg["cell"] == "lstm" else 3
    H, d, r = cfg["H"], cfg["d"], cfg["r"]
    ranks = [1] + [r] * (d - 1) + [1]
    total = 0
    hh = chain_flops(*tt_shape(H, H, d, G), ranks)
    for l in range(cfg["L"]):
        n_in = cfg["I"] if l == 0 else H
        total += chain_flops(*tt_shape(n_in, H, d, G), ranks) + hh + 2 * G * H + 12 * H
    return total, hh


def make_input(cfg, B, device="cpu"):
    g = torch.Generator().manual_seed(cfg["seed"])
    x = torch.rand(B, cfg["T"], cfg["I"], generator=g)
    if cfg["inp"].startswith("digits"):
        x = (x - 0.1307) / 0.3081                      # reference digit_classification/utils.py:9-10
    if cfg["inp"] == "digits_perm":
        np.random.seed(cfg["seed"])
        perm = torch.from_numpy(np.random.permutation(cfg["T"]))   # pmnist_test.py:92,132,154
        x = x[:, perm, :].contiguous()
    return x.to(device)


def upstream(cfg, out, hT):
    """Scalar whose gradient is the config's upstream gradient pattern."""
    if cfg["grad"] == "out_last":
        return out[:, -1, :].sum()
    if cfg["grad"] == "hT":
        return hT.sum()
    return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def run_cpu_oracle(cfg, B, T, steps, warmup):
    """Time the CPU oracle (restatement of the reference's PyTorch path) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ttrnn_oracle as oracle                      # bench: cpu_baseline / reference arm only
    torch.set_num_threads(os.cpu_count() or 1)
    c = dict(cfg)
    c["T"] = T
    layers = oracle.random_layers(cfg["cell"], cfg["I"], cfg["H"], cfg["L"], cfg["d"], cfg["r"], bias=True,
                                  seed=cfg["seed"], requires_grad=(cfg["mode"] != "fwd"))
    x = make_input(c, B)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if cfg["mode"] == "fwd":
            with torch.no_grad():
                if cfg["cell"] == "lstm":
                    oracle.lstm_forward(layers, x)
                else:
                    oracle.gru_forward(layers, x)
        else:
            for p in oracle.flat_params(layers):
                p.grad = None
            if cfg["cell"] == "lstm":
                out, (h, _) = oracle.lstm_forward(layers, x)
            else:
                out, h = oracle.gru_forward(layers, x)
            loss = upstream(cfg, out, h)
            if loss is None:
                gd = torch.Generator().manual_seed(5)
                loss = (out * torch.rand(out.shape, generator=gd)).sum()
            loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return B * T / sec, sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--seq-len", type=int, default=0, help="override T (profiling only; not a bench number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["B"] = args.batch
    if args.seq_len:
        cfg["T"] = args.seq_len
        cfg["name"] += " [T overridden to %d: profiling run]" % args.seq_len
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config_json = {"workload": cfg["name"], "cell": cfg["cell"], "input_size": cfg["I"], "hidden_size": cfg["H"],
                   "num_layers": cfg["L"], "n_cores": cfg["d"], "tt_rank": cfg["r"], "batch_per_gpu": cfg["B"],
                   "global_batch": cfg["B"] * world, "seq_len": cfg["T"], "mode": cfg["mode"],
                   "upstream_grad": cfg["grad"], "parallelism": "dp%d (batch-sharded replicas)" % world,
                   "l2": "flushed between timed steps (256 MiB write)"}

    # ---------------- reference arm: the CPU path on the host cores --------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        Bs, Ts = CPU_SAMPLE_B[args.config], CPU_SAMPLE_T[args.config]
        val, sec = run_cpu_oracle(cfg, Bs, Ts, args.steps, args.warmup)
        sample = "batch %d x T %d of the workload per step (T %s); oracle port of the reference PyTorch path" % (
            Bs, Ts, "as configured" if Ts == cfg["T"] else "reduced from %d" % cfg["T"])
        line = {"impl": "reference", "metric": "TT-RNN cell-steps/sec (batch x T), %s" % cfg["mode"], "value": val,
                "unit": "cell-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_json,
                "cpu_baseline": {"value": val, "unit": "cell-steps/s", "cores": torch.get_num_threads(),
                                 "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- our arm -----------------------------------------------------------------
    import tensorized_rnn_b200 as tr
    from tensorized_rnn_b200 import _lib
    from tensorized_rnn_b200.dist import allreduce_gradients
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    torch.manual_seed(cfg["seed"])
    cls = tr.TTLSTM if cfg["cell"] == "lstm" else tr.TTGRU
    with redirect_stdout(io.StringIO()):
        model = cls(cfg["I"], cfg["H"], cfg["L"], torch.device("cpu"), n_cores=cfg["d"], tt_rank=cfg["r"]).to(dev)
    params = [p for p in model.parameters()]
    B, T, H = cfg["B"], cfg["T"], cfg["H"]
    c_rank = dict(cfg)
    c_rank["seed"] = cfg["seed"] + rank               # every rank owns different sequences
    x_host = make_input(c_rank, B).pin_memory()
    x_dev = x_host.to(dev)
    train = cfg["mode"] != "fwd"
    dense_dout = None
    if cfg["grad"] == "dense":
        dense_dout = torch.rand(B, T, H, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one_pass(x):
        if not train:
            with torch.no_grad():
                res = model(x)
            return res[1][0] if cfg["cell"] == "lstm" else res[1]
        for p in params:
            p.grad = None
        res = model(x)
        out = res[0]
        hT = res[1][0] if cfg["cell"] == "lstm" else res[1]
        if dense_dout is not None:
            out.backward(dense_dout)
            result = hT
        else:
            loss = upstream(cfg, out, hT)
            loss.backward()
            result = loss
        if dist is not None:
            allreduce_gradients(params)                 # one flat NCCL all-reduce of every TT-core / bias gradient
        return result

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # FP32 peak: FFMA probe, best of 5 (burst figure for a kernel timed alone)
    sink = torch.zeros(4, device=dev)
    flops = C.c_double(0)
    stream = torch.cuda.current_stream().cuda_stream
    peak = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        _lib.check(lib.ttrnn_ffma_probe(4000, sink.data_ptr(), C.byref(flops), stream), "ttrnn_ffma_probe")
        e1.record()
        torch.cuda.synchronize()
        peak = max(peak, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    for _ in range(args.warmup):
        one_pass(x_dev)
    barrier()

    # ---- value: device-resident, per-step CUDA events, L2 flushed between steps -------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.ttrnn_launch_count(1)
    lib.ttrnn_kernel_timing(1)
    ev = []
    barrier()
    for _ in range(args.steps):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        one_pass(x_dev)
        e1.record()
        ev.append((e0, e1))
    barrier()
    lib.ttrnn_kernel_timing(0)
    launches = int(lib.ttrnn_launch_count(0))
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    kms = (C.c_double * 4)(0, 0, 0, 0)
    kcnt = (C.c_int64 * 4)(0, 0, 0, 0)
    lib.ttrnn_kernel_times(kms, kcnt)

    # ---- e2e: pinned host input -> H2D -> module API -> D2H of the result, wall clock ----------
    x_stage = torch.empty_like(x_dev)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        x_stage.copy_(x_host, non_blocking=True)
        res = one_pass(x_stage)
        host = res.detach().to("cpu")                   # D2H read of the step's result (loss / final state)
        d2h = host.numel() * host.element_size()
    barrier()
    e2e_s = time.perf_counter() - t0

    tot = torch.tensor([total_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(tot[0]), float(tot[1])

    if rank == 0:
        units = B * T * world * args.steps
        value = units / (total_ms * 1e-3)
        fwd_flops, hh_flops = algorithmic_flops(cfg)
        mult = 3 if train else 1
        names = ["k_ttlinear_fwd", "k_rnn_fwd", "k_rnn_bwd", "k_ttlinear_bwd"]
        share = {names[i]: {"ms_per_step": kms[i] / args.steps, "launches_per_step": kcnt[i] / args.steps}
                 for i in range(4)}
        dom = max(range(4), key=lambda i: kms[i])
        # algorithmic FLOPs of the dominant kernel per step (recomputed forward work is NOT credited)
        G = 4 if cfg["cell"] == "lstm" else 3
        gate = 2 * G * H + 12 * H
        ih_flops = fwd_flops - cfg["L"] * (hh_flops + gate)
        per_seqstep = {0: ih_flops * (1 if train else 1), 1: cfg["L"] * (hh_flops + gate),
                       2: 2 * cfg["L"] * (hh_flops + gate), 3: 2 * ih_flops}[dom]
        dom_flops = per_seqstep * B * T
        dom_ms = kms[dom] / args.steps
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        whole = mult * fwd_flops * B * T * world * args.steps / (total_ms * 1e-3) / 1e12 / world
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                ent = json.load(f).get(cfg["name"])
            if ent and ent["kernel"] == names[dom].replace("k_ttlinear", "k_ttlinear"):
                traffic = ent["dram_bytes_per_launch"]
        except Exception:
            traffic = None
        # contraction order the engine chose for each layer's batched ih projection (0 = TT chain, 1 = dense
        # W_ih formed once per call, 2 = rank-one input): the roofline figures above always credit the
        # reference's core-by-core sweep; where the dense order is cheaper the kernels execute
        # FEWER multiply-adds than credited, so both costs are reported
        routes = []
        desc = model.spec().desc(B, T)
        for l in range(cfg["L"]):
            cm, dm = C.c_int64(0), C.c_int64(0)
            rt = int(lib.ttrnn_rnn_ih_route(C.byref(desc), l, C.byref(cm), C.byref(dm)))
            routes.append({"layer": l, "route": {0: "tt_chain", 1: "dense", 2: "rank_one"}.get(rt, str(rt)),
                           "chain_macs_per_row": cm.value, "dense_macs_per_row": dm.value})
        line = {
            "metric": "TT-RNN cell-steps/sec (batch x T), %s" % cfg["mode"],
            "value": value, "unit": "cell-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (random-init TT cores, %s input)" % cfg["inp"],
            "config": config_json,
            "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "cell-steps/s",
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp32_ffma", "kernel": names[dom], "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_unit": "bytes per launch (ncu dram read+write, profiles/)",
                         "peak_source": "ttrnn_ffma_probe measured in this run (MEASURED_PEAKS.json has no FP32 entry)",
                         "algorithmic_flops_per_seqstep": per_seqstep,
                         "whole_step_tflops_per_gpu": whole, "whole_step_frac": whole / peak if peak else None,
                         "fwd_flops_per_seqstep": fwd_flops, "kernels": share, "ih_projection": routes},
        }
        if world == 1 and not args.no_cpu_baseline:
            Bs, Ts = CPU_SAMPLE_B[args.config], CPU_SAMPLE_T[args.config]
            val, sec = run_cpu_oracle(cfg, Bs, Ts, 1, 1 if sec_budget_ok(args.config) else 0)
            line["cpu_baseline"] = {"value": val, "unit": "cell-steps/s", "cores": torch.get_num_threads(),
                                    "kind": "port", "seconds": sec,
                                    "sample": "batch %d x T %d of the workload, 1 pass, oracle port of the "
                                              "reference PyTorch path" % (Bs, Ts)}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def sec_budget_ok(config):
    return True


if __name__ == "__main__":
    sys.exit(main())
