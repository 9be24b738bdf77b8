#!/usr/bin/env python
"""bench.py - headline benchmark of the DPF-Nets hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU

Metric (BASELINE.json): decoder points/s, train forward+backward, on the chair generation config
(configs/generation/chair.yaml: 21 triples = 63 coupling layers, F=64, G=128), synthetic batch
32 x 2048 points per GPU, random-init weights.  One "step" = LocalCondRNVPDecoder(p, g, 'inverse')
in train mode + PointFlowNLL + backward to every decoder parameter and g (the
optimizer is not part of this metric).  N > 1: batch-sharded (weak scaling), NCCL all-reduce of
the gradient arena inside the step.  Also reported: sampling points/s, Chamfer pair-evals/s,
roofline of the dominant kernel, and the oracle port timed on the host cores (cpu_baseline).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs (reference arm, cpu_baseline) run on rank 0 only and
# must see all host cores, and the thread pools are sized when torch is imported - so this comes first.
if int(os.environ.get("RANK", "0")) == 0:
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import torch  # noqa: E402

N_FLOWS, F_WIDTH, G_LATENT = 21, 64, 128
BASE_LOGVAR = -3.6990                    # configs/generation/chair.yaml:51 p_decoder_base_var
FLOP_PER_POINT_LAYER_FWD = 17152         # 2 branches x (2k64 + 2*64*64 + 2*64w), k+w=3
KERNEL_CLASSES = ["film_fwd", "moments", "fwd_stats", "fwd_apply", "bwd_p1", "bwd_p2", "bwd_final", "film_bwd"]
# algorithmic (non-recompute) GEMM FLOP per point per layer each kernel class is responsible for
# dram__bytes_read.sum + dram__bytes_write.sum per launch of each per-layer kernel, from the `ncu --set full` captures
# of this workload kept under profiles/ (r01_ncu_coupling_*_summary.txt); inputs of a layer mostly hit in the 126 MB L2


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each per-layer kernel class, from the newest
    profiles/r*_ncu_traffic.json (written by tools/ncu_summary.py from an `ncu --set full` capture of THIS workload;
    inputs of a layer mostly hit in the 126 MB L2)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return {}, None
    with open(files[-1]) as f:
        return {k: v for k, v in json.load(f).items() if not k.startswith("_")}, os.path.relpath(files[-1], ROOT)
ALGO_FLOP = {"fwd_stats": 0, "fwd_apply": FLOP_PER_POINT_LAYER_FWD, "bwd_p1": 2 * 2 * 64 * 3,
             "bwd_p2": 2 * FLOP_PER_POINT_LAYER_FWD - 2 * 2 * 64 * 3}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DPF_PRECISION", "auto"))
    ap.add_argument("--batch", type=int, default=32, help="shapes per GPU")
    ap.add_argument("--points", type=int, default=2048)
    ap.add_argument("--no-extras", action="store_true", help="skip sampling / Chamfer / cpu_baseline legs")
    ap.add_argument("--sweep-clouds", type=int, default=1000,
                    help="clouds per set of the row-sharded evaluation sweep in the extras (BASELINE config 5: 1000; 0 = skip)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph replay per step")
    ap.add_argument("--side-stream", action="store_true", help="run on a non-default CUDA stream (A/B tests)")
    ap.add_argument("--lib-option", action="append", default=[], metavar="K=V", help="dpf_set_option(K, V) before the run (A/B tests)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def synth_inputs(B, N, G, rank, pin=False):
    gen = torch.Generator().manual_seed(1234 + rank)
    p = torch.rand((B, 3, N), generator=gen) - 0.5
    g = torch.randn((B, G), generator=gen)
    if pin:
        p, g = p.pin_memory(), g.pin_memory()
    return p, g


class ClockSampler:
    """nvidia-smi sampling during the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port (CPU restatement of the reference algorithm)
# --------------------------------------------------------------------------------------------
def oracle_train_step_factory(B, N, rank=0, device="cpu"):
    """The oracle decoder (oracle/flow_oracle.py: torch restatement of the reference modules, reference-faithful random
    init from the oracle's own initialiser - no product code on this arm) -> step() running fwd + PointFlowNLL + bwd."""
    from oracle import flow_oracle as fo
    layers, leaves = fo.init_decoder_layers(N_FLOWS, G_LATENT, F=F_WIDTH, weight_std=0.01, seed=0, device=device)
    p, g = synth_inputs(B, N, G_LATENT, rank)
    p, g = p.to(device), g.to(device)
    g.requires_grad_(True)
    base_mu, base_lv = torch.zeros_like(p), torch.full_like(p, BASE_LOGVAR)

    def step():
        for t in leaves:
            t.grad = None
        g.grad = None
        ps, mus, lvs = fo.decoder_forward(layers, p, g, "inverse", training=True)
        nll = fo.point_flow_nll(ps + [p], [base_mu] + mus, [base_lv] + lvs)
        nll.backward()
        return float(nll.detach())
    return step


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    """Reference arm: the reference algorithm (oracle port, torch CPU fp32) on the box's host cores.  Under torchrun
    only rank 0 runs it - ONE process, one batch_per_gpu x points batch per step whatever --gpus says - and the
    config says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, N = args.batch, args.points
    step = oracle_train_step_factory(B, N)
    t0 = time.perf_counter()
    step()
    est = time.perf_counter() - t0
    sample = "full batch %dx%d per step" % (B, N)
    if est * (args.steps + args.warmup) > 240.0 and B > 8:
        B = 8
        step = oracle_train_step_factory(B, N)
        sample = "bounded sample: %d of %d shapes x %d points per step" % (B, args.batch, N)
    ts = time_cpu(step, args.steps, args.warmup)
    ms = 1e3 * sum(ts) / len(ts)
    val = B * N / (ms * 1e-3)
    cfg = workload_config(args, "fp32")
    cfg.update(global_batch=B, parallelism="host CPU, 1 process x %d threads (rank 0 only; --gpus %d does not add work)" % (torch.get_num_threads(), args.gpus))
    line = {
        "impl": "reference", "metric": "decoder points/s (train fwd+bwd)", "value": val, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, precision):
    return {"workload": "chair generation config (configs/generation/chair.yaml): point decoder, 63 coupling "
                        "layers F=64 G=128, train-mode inverse pass + PointFlowNLL + backward",
            "batch_per_gpu": args.batch, "points": args.points, "global_batch": args.batch * args.gpus,
            "precision": precision, "parallelism": "dp%d" % args.gpus,
            "l2": "256 MiB flush write between timed steps (untimed)",
            "launch": "one CUDA graph replay per step (captured from the public module API), eager fallback",
            "optimizer": "not part of the metric (SURVEY.md 8d)"}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from dpf_nets_b200 import _lib
    from dpf_nets_b200.lib.networks.decoders import LocalCondRNVPDecoder, prepend
    from dpf_nets_b200.lib.networks.losses import PointFlowNLL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    _lib.check(lib.dpf_device_check(), "dpf_device_check")
    for kv in args.lib_option:      # A/B switches of the library (include/dpfnets_b200.h: dpf_set_option)
        k, v = kv.split("=")
        _lib.check(lib.dpf_set_option(int(k), int(v)), "dpf_set_option")
    pk = peaks()

    precision = args.precision
    torch.manual_seed(0)
    model = LocalCondRNVPDecoder(N_FLOWS, F_WIDTH, G_LATENT).to(dev)
    model.precision = precision
    model.train()
    if precision == "auto":   # what the module resolves 'auto' to in train mode (plain bf16 in eval)
        from dpf_nets_b200.lib.networks._flowfn import resolve_precision
        precision = resolve_precision(model)
    B, N = args.batch, args.points
    p_host, g_host = synth_inputs(B, N, G_LATENT, rank, pin=True)
    p_dev, g_dev = p_host.to(dev), g_host.to(dev).requires_grad_(True)
    base_mu, base_lv = torch.zeros_like(p_dev), torch.full_like(p_dev, BASE_LOGVAR)
    crit = PointFlowNLL()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def compute(p, g):
        """module call + PointFlowNLL + backward (no collective)"""
        model.arena.grad = None
        g.grad = None
        ps, mus, lvs = model(p, g, mode="inverse")
        nll = crit(prepend(None, ps)[1:] + [p], prepend(base_mu, mus), prepend(base_lv, lvs))
        nll.backward()
        return nll

    def step(p, g):
        nll = compute(p, g)
        if world > 1:
            dist.all_reduce(model.arena.grad, op=dist.ReduceOp.AVG)
        return nll

    def launches():
        n = ctypes.c_longlong(0)
        lib.dpf_launch_count(ctypes.byref(n))
        return n.value

    for _ in range(args.warmup):
        step(p_dev, g_dev)
    torch.cuda.synchronize()

    # The step is launch-dense (~195 kernels of ~25 us): like the product's `--cuda_graph` training step
    # (lib/networks/_graphstep.py) the bench replays module call + loss + backward as ONE CUDA graph captured once from the
    # public API on static input buffers; the gradient all-reduce (N > 1) stays an eager NCCL call after the replay.
    # `--no-graph` (or a failed capture) times the eager launches instead; `launch_mode` in the JSON line says which.
    p_in, g_in = p_dev.clone(), g_dev.detach().clone().requires_grad_(True)
    graph, static_loss, static_grad, launches_per_step, launch_mode = None, None, None, None, "eager launches"
    if not args.no_graph:
        try:
            model.arena.grad = None
            g_in.grad = None
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            lc0 = launches()
            with torch.cuda.graph(graph):
                static_loss = compute(p_in, g_in).detach()
            launches_per_step = launches() - lc0
            static_grad = model.arena.grad          # lives in the graph's memory pool: every replay rewrites it in place
            graph.replay()
            torch.cuda.synchronize()
            launch_mode = "one CUDA graph replay per step (module call + PointFlowNLL + backward captured once from the public API)"
        except Exception as exc:     # never fatal for the headline line
            graph, launch_mode = None, "eager launches (graph capture failed: %s: %s)" % (type(exc).__name__, exc)
            torch.cuda.synchronize()

    def run_step():
        if graph is None:
            return step(p_in, g_in)
        graph.replay()
        if world > 1:
            dist.all_reduce(static_grad, op=dist.ReduceOp.AVG)
        return static_loss

    for _ in range(args.warmup):
        run_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- timed region: device-resident inputs ----
    sampler = ClockSampler(local)
    sampler.start()

    def timed(prepare):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        val = 0.0
        torch.cuda.synchronize()
        for a, b in evs:
            flush.zero_()
            a.record()
            val = prepare()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return sum(a.elapsed_time(b) for a, b in evs), val

    l0 = launches()
    total_ms, _ = timed(run_step)
    l1 = launches()
    n_launches = (l1 - l0) if graph is None else launches_per_step * args.steps

    # ---- e2e: pinned host inputs -> H2D -> step -> D2H of the loss, every step ----
    def e2e_step():
        p_in.copy_(p_host, non_blocking=True)
        with torch.no_grad():
            g_in.copy_(g_host, non_blocking=True)
        return float(run_step().item())
    e2e_ms, loss_val = timed(e2e_step)
    # the same two measurements with eager launches (host launch overhead and jitter exposed), for the record
    if graph is not None:
        eager_ms, _ = timed(lambda: step(p_dev, g_dev))
    else:
        eager_ms = total_ms
    clocks = sampler.stop()
    t = torch.tensor([total_ms, e2e_ms, eager_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, eager_ms = float(t[0]), float(t[1]), float(t[2])
    ms_per_step = total_ms / args.steps
    value = world * B * N / (ms_per_step * 1e-3)
    e2e_val = world * B * N / (e2e_ms / args.steps * 1e-3)

    # ---- per-kernel-class timing (separate pass, CUDA events on the launching stream) ----
    lib.dpf_profile_enable(1)
    prof_steps = max(2, min(args.steps, 5))
    for _ in range(prof_steps):
        flush.zero_()
        step(p_dev, g_dev)
    torch.cuda.synchronize()
    ms = (ctypes.c_double * 8)()
    cnt = (ctypes.c_longlong * 8)()
    lib.dpf_profile_collect(ms, cnt, 8)
    lib.dpf_profile_enable(0)
    classes = {k: {"ms_per_step": ms[i] / prof_steps, "launches_per_step": cnt[i] / prof_steps,
                   "us_per_launch": (1e3 * ms[i] / cnt[i]) if cnt[i] else None} for i, k in enumerate(KERNEL_CLASSES)}
    traffic, traffic_src = ncu_traffic()
    dom = max(("fwd_stats", "fwd_apply", "bwd_p1", "bwd_p2"), key=lambda k: classes[k]["ms_per_step"])
    dom_us = classes[dom]["us_per_launch"] or float("inf")
    tensor_peak = pk["bf16_tflops_sustained"]
    achieved = ALGO_FLOP[dom] * B * N / (dom_us * 1e-6) / 1e12
    whole = 3 * FLOP_PER_POINT_LAYER_FWD * 3 * N_FLOWS * B * N / (ms_per_step * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak,
                "traffic": traffic.get(dom) if (B, N) == (32, 2048) else None,
                "traffic_source": ("ncu --set full capture of this kernel on this workload: %s" % traffic_src) if traffic_src else None,
                "peak_source": pk["source"] + " (sustained bf16)",
                "algorithmic_flop_per_launch": ALGO_FLOP[dom] * B * N,
                "whole_step": {"achieved": whole, "frac": whole / tensor_peak,
                               "flop_per_point": 3 * FLOP_PER_POINT_LAYER_FWD * 3 * N_FLOWS},
                "kernel_classes": classes}

    line = {
        "metric": "decoder points/s (train fwd+bwd)", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if precision.startswith("bf16") else "f32",
        "data": "synthetic", "config": workload_config(args, precision),
        "e2e": {"value": e2e_val, "unit": "points/s", "h2d_bytes_per_step": p_host.numel() * 4 + g_host.numel() * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps, "loss": loss_val},
        "launch_mode": launch_mode, "eager_ms_per_step": eager_ms / args.steps,
        "gpu_launches": n_launches, "clocks": clocks, "roofline": roofline,
    }

    if not args.no_extras:
        extra = {}
        if args.sweep_clouds > 0:                       # every rank takes part in the sharded sweep
            extra["eval_sweep"] = eval_sweep(dev, args.sweep_clouds, N, world, rank, flush)
        # whole-model training steps of the three BASELINE model families, batch-sharded over the ranks (gradient
        # averaging: dist.GradSync - decoder arena all-reduce overlapped with the rest of the backward)
        extra["full_model_step"] = full_model_step(dev, B, N, precision, flush, "generation/chair", world, graphed=True)
        extra["full_model_step_ae_all_original"] = full_model_step(dev, B, N, precision, flush, "autoencoding/all_original", world, graphed=True)
        extra["full_model_step_svr_all"] = full_model_step(dev, B, N, precision, flush, "svr/all", world, graphed=True)
        if world == 1:
            # single-GPU legs and the CPU baseline only in the N = 1 run (under torchrun the other ranks would idle in a
            # barrier while rank 0 works, and the driver's N = 1 line already carries them)
            extra.update(extras(model, dev, B, N, pk, flush))
            extra["decoder_b64"] = decoder_step_b64(dev, N, args.precision, flush)
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cstep = oracle_train_step_factory(B, N)
            ts = time_cpu(cstep, 2, 1)
            cms = 1e3 * sum(ts) / len(ts)
            line["cpu_baseline"] = {"value": B * N / (cms * 1e-3), "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "full batch %dx%d, 1 warm-up + 2 timed steps of the oracle port "
                                              "(torch CPU fp32 restatement of the reference modules)" % (B, N),
                                    "ms_per_step": cms}
        if rank == 0:
            line["extra"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def decoder_step_b64(dev, N, precision, flush, steps=5, warmup=3):
    """The YAML default batch (batch_size: 64, configs/generation/*.yaml): 1024 tiles do not fit the merged forward's
    TMEM-resident form (592 tile slots), so the forward runs as statistics + apply launches with a recompute."""
    from dpf_nets_b200.lib.networks.decoders import LocalCondRNVPDecoder, prepend
    from dpf_nets_b200.lib.networks.losses import PointFlowNLL
    B = 64
    torch.manual_seed(0)
    model = LocalCondRNVPDecoder(N_FLOWS, F_WIDTH, G_LATENT).to(dev).train()
    model.precision = precision
    p, g = synth_inputs(B, N, G_LATENT, 0)
    p, g = p.to(dev), g.to(dev).requires_grad_(True)
    base_mu, base_lv = torch.zeros_like(p), torch.full_like(p, BASE_LOGVAR)
    crit = PointFlowNLL()

    def step():
        model.arena.grad = None
        g.grad = None
        ps, mus, lvs = model(p, g, mode="inverse")
        crit(prepend(None, ps)[1:] + [p], prepend(base_mu, mus), prepend(base_lv, lvs)).backward()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    return {"value": B * N / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "batch": B, "points": N,
            "forward_form": "two launches per layer (statistics + apply): 1024 tiles exceed the merged forward's 592 resident tile slots"}


def measured_pipe_ceilings(dev):
    """MUFU (ex2.approx) results/s and packed-fp32 FMA lanes/s of this GPU, measured with dpf_throughput_probe."""
    from dpf_nets_b200 import _lib
    lib = _lib.lib()
    scratch = torch.zeros(4, device=dev)
    out = {}
    for name, which in (("mufu_ex2_per_s", 0), ("fp32_fma_lanes_per_s", 1)):
        n = ctypes.c_longlong(0)
        best = 0.0
        for it in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.dpf_throughput_probe(which, 148 * 8, 4096, _lib.ptr(scratch), ctypes.byref(n), _lib.stream_ptr(dev)), "dpf_throughput_probe")
            b.record()
            torch.cuda.synchronize()
            if it:
                best = max(best, n.value / (a.elapsed_time(b) * 1e-3))
        out[name] = best
    return out


def extras(model, dev, B, N, pk, flush):
    """Secondary metrics BASELINE.json names: sampling points/s and Chamfer pair-evals/s."""
    from dpf_nets_b200.ops import pairwise_cd
    out = {}
    model.eval()
    gen = torch.Generator().manual_seed(7)
    z = torch.randn((B, 3, N), generator=gen).to(dev)
    g = torch.randn((B, G_LATENT), generator=gen).to(dev)
    with torch.no_grad():
        for _ in range(3):
            model(z, g, mode="direct")
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); model(z, g, mode="direct"); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    samp = B * N / (ms * 1e-3)
    out["sampling"] = {"value": samp, "unit": "points/s", "ms_per_pass": ms, "precision": "bf16 ('auto' = plain bf16 in eval mode, 2e-2 gate)",
                       "tensor_frac": samp * FLOP_PER_POINT_LAYER_FWD * 3 * N_FLOWS / 1e12 / pk["bf16_tflops_sustained"]}
    model.train()
    out["encoder_eval"] = encoder_eval(dev, B, N, flush)
    out["encoder_train"] = encoder_train(dev, B, N, flush)
    S = 256
    A = (torch.rand((S, N, 3), generator=gen) - 0.5).to(dev)
    Bc = (torch.rand((S, N, 3), generator=gen) - 0.5).to(dev)
    pairwise_cd(A, Bc)
    ts = []
    for _ in range(3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pairwise_cd(A, Bc); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    sec = sum(ts) / len(ts) * 1e-3
    pairs = S * S / sec
    # Bound: the FP32 (FMA) pipe.  One evaluation of a point pair = 6 fp32 lane-operations (3 sub, 1 mul, 2 fma; the packed
    # FADD2 / FMUL2 / FFMA2 halve the ISSUE slots, not the pipe work; the mins run on the ALU pipe); SURVEY 8d counts the
    # algorithmic N^2 evaluations per cloud pair (one evaluation serves both Chamfer directions in the shipped kernel).
    ceil = measured_pipe_ceilings(dev)
    fp32_nominal = 148 * 128 * 1.965e9 / 6.0
    fp32_measured = ceil["fp32_fma_lanes_per_s"] / 6.0
    out["chamfer"] = {"value": pairs, "unit": "cloud-pair CD evals/s", "clouds": "%dx%d of %d points" % (S, S, N),
                      "point_pair_evals_per_s": pairs * N * N,
                      "hbm": {"achieved_gbs": pairs * (2 * N * 12 + 4) / 1e9, "peak_gbs": pk["hbm_gbs"],
                              "frac": pairs * (2 * N * 12 + 4) / 1e9 / pk["hbm_gbs"]},
                      "fp32_pipe": {"achieved_dist_evals_per_s": pairs * N * N, "evals_per_cloud_pair": N * N,
                                    "bound_nominal": fp32_nominal, "bound_measured": fp32_measured,
                                    "measured_fp32_fma_lanes_per_s": ceil["fp32_fma_lanes_per_s"],
                                    "frac": pairs * N * N / fp32_measured, "frac_of_nominal": pairs * N * N / fp32_nominal}}
    # approximate EMD, fused all-pairs cost (match never materialised): MUFU(ex2)-bound, 27 sweeps x N^2 exps per pair
    from dpf_nets_b200.ops import pairwise_emd
    Se = 64
    pairwise_emd(A[:8].contiguous(), Bc[:8].contiguous())
    ts = []
    for _ in range(2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pairwise_emd(A[:Se].contiguous(), Bc[:Se].contiguous()); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    sec = sum(ts) / len(ts) * 1e-3
    epairs = Se * Se / sec
    mufu = ceil["mufu_ex2_per_s"]                           # measured on this GPU (dpf_throughput_probe), not assumed
    out["emd"] = {"value": epairs, "unit": "cloud-pair approximate-EMD evals/s", "clouds": "%dx%d of %d points" % (Se, Se, N),
                  "exp_per_pair": 27 * N * N, "mufu": {"achieved_ex2_per_s": epairs * 27 * N * N, "bound_measured": mufu,
                                                        "mufu_results_per_clk_per_sm": mufu / (148 * 1.965e9),
                                                        "frac": epairs * 27 * N * N / mufu,
                                                        "executed_mufu_per_pair": 36 * N * N,   # + 9 N^2 rsqrt of the fused cost
                                                        "frac_executed": epairs * 36 * N * N / mufu}}
    return out


def eval_sweep(dev, S, N, world, rank, flush):
    """BASELINE config 5: gg, tt, gt Chamfer matrices of S generated vs S reference clouds of N points, rows
    sharded across the ranks (interleaved; gg / tt upper triangle only), ONE collective per matrix, then
    COV / MMD / 1-NNA on every rank (lib/networks/evaluating.py:245-253 -> utils.pairwise_CD / COV / MMD / KNN).
    Timed on the device, max over ranks; all ranks take part."""
    import torch.distributed as dist
    from dpf_nets_b200.lib.networks.utils import pairwise_CD
    from dpf_nets_b200.ops.metrics import cd_scores
    gen = torch.Generator().manual_seed(4321)          # every rank holds both sets (24.6 MB each at 1000 x 2048)
    G = (torch.rand((S, N, 3), generator=gen) - 0.5).to(dev)
    R = (torch.rand((S, N, 3), generator=gen) - 0.5).to(dev)
    small = G[:8].contiguous()
    pairwise_CD(small, small)                          # warm-up (kernel attributes, NCCL channels)
    pairwise_CD(small, R[:8].contiguous())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    flush.zero_()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    gg = pairwise_CD(G, G)
    tt = pairwise_CD(R, R)
    gt = pairwise_CD(G, R)
    b.record()
    scores = cd_scores(gg, gt, tt)                     # one reduction launch on the device-resident matrices
    c.record()
    torch.cuda.synchronize()
    cov, mmd, nna = scores.tolist()
    t = torch.tensor([a.elapsed_time(b), a.elapsed_time(c)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_mat, ms_all = float(t[0]), float(t[1])
    evaluated = S * (S + 1) + S * S                    # cloud pairs actually evaluated (two triangles + one full matrix)
    return {"clouds": "%d generated x %d reference of %d points" % (S, S, N), "n_gpus": world,
            "ms_matrices": ms_mat, "ms_with_cov_mmd_1nna": ms_all,
            "value": evaluated / (ms_mat * 1e-3), "unit": "cloud-pair CD evals/s (pairs evaluated; symmetric halves of gg/tt skipped)",
            "matrix_entries_per_s": 3.0 * S * S / (ms_mat * 1e-3),
            "point_pair_evals_per_s": evaluated / (ms_mat * 1e-3) * N * N,
            "scores": {"COV": cov, "MMD": mmd, "1-NNA": nna}, "sharding": "interleaved rows, one all-reduce(sum) of disjoint row blocks per matrix"}


def encoder_eval(dev, B, N, flush):
    """PointNet encoder + max-pool in eval mode: fused tcgen05 kernel vs the library path."""
    from dpf_nets_b200.lib.networks.encoders import PointNetCloudEncoder
    torch.manual_seed(0)
    enc = PointNetCloudEncoder(3, 64, [128, 256, 512]).to(dev).eval()
    x = (torch.rand((B, 3, N), generator=torch.Generator().manual_seed(3)) - 0.5).to(dev)
    res = {}
    with torch.no_grad():
        for name, prec in (("fused", "auto"), ("library_path", "fp32")):
            enc.precision = prec
            for _ in range(3):
                enc.global_features(x)
            ts = []
            for _ in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); enc.global_features(x); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            res[name + "_ms"] = sum(ts) / len(ts)
    ms = res["fused_ms"]
    flop = 2.0 * (3 * 64 + 64 * 128 + 128 * 256 + 256 * 512) * B * N
    res.update({"value": B * N / (ms * 1e-3), "unit": "points/s", "tflops": flop / (ms * 1e-3) / 1e12})
    return res


def encoder_train(dev, B, N, flush):
    """PointNet encoder + max-pool in TRAIN mode, forward + backward: the whole encoder as one autograd
    function over the library's own tcgen05 kernels vs the library path of the same module (bmm + ATen batch-norm + clamp)."""
    from dpf_nets_b200.lib.networks.encoders import PointNetCloudEncoder
    torch.manual_seed(0)
    enc = PointNetCloudEncoder(3, 64, [128, 256, 512]).to(dev).train()
    x = (torch.rand((B, 3, N), generator=torch.Generator().manual_seed(3)) - 0.5).to(dev)
    cot = torch.randn((B, 512), generator=torch.Generator().manual_seed(4)).to(dev)
    res = {}
    for name, prec in (("fused", "auto"), ("library_path", "fp32")):
        enc.precision = prec

        def step():
            enc.zero_grad()
            (enc.global_features(x) * cot).sum().backward()
        for _ in range(3):
            step()
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[name + "_ms"] = sum(ts) / len(ts)
    res.update({"value": B * N / (res["fused_ms"] * 1e-3), "unit": "points/s", "what": "forward + backward, eager launches"})
    return res


def full_model_step(dev, B, N, precision, flush, config_name, world, graphed=False, steps=5, warmup=3):
    """Whole training step of one of the BASELINE model families ('also report whole-model step'):
    encoders + latent flows + priors + point decoder + VAE loss + backward + gradient averaging across the ranks
    (batch sharding, B shapes per rank) + AMSGrad Adam.  Timed on every rank, max over ranks.
      generation/chair           BASELINE config 2 (G = 128)
      autoencoding/all_original  BASELINE config 3 (G = 512), 'batch-sharded training with gradient allreduce'
      svr/all                    BASELINE config 4 (ResNet-18 image encoder + G = 512), 224 x 224 synthetic images"""
    import torch.distributed as dist
    from dpf_nets_b200 import configs
    from dpf_nets_b200 import dist as ddist
    from dpf_nets_b200.lib.networks.losses import Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss
    from dpf_nets_b200.lib.networks.models import Local_Cond_RNVP_MC_Global_RNVP_VAE, Local_Cond_RNVP_MC_Global_RNVP_VAE_IC
    from dpf_nets_b200.lib.networks.optimizers import Adam
    config = configs.load(config_name)
    ic = config["train_mode"] == "p_rnvp_mc_g_rnvp_vae_ic"
    rank = int(os.environ.get("RANK", "0"))
    torch.manual_seed(0)
    model = (Local_Cond_RNVP_MC_Global_RNVP_VAE_IC if ic else Local_Cond_RNVP_MC_Global_RNVP_VAE)(**config).to(dev)
    model.pc_decoder.precision = precision
    model.train()
    crit = Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(**config).to(dev)
    opt = Adam(model.parameters(), lr=config["max_lr"], weight_decay=config["wd"], betas=(config["beta1"], config["max_beta2"]),
               amsgrad=True)
    sync = ddist.GradSync(model)
    gen = torch.Generator().manual_seed(99 + rank)
    host = [(torch.rand((B, 3, N), generator=gen) - 0.5).pin_memory() for _ in range(2)]
    if ic:
        host.append(torch.randn((B, 4, 224, 224), generator=gen).pin_memory())

    def step():
        inp = [t.to(dev, non_blocking=True) for t in host]
        out = model(*inp)
        loss, pnll, gnll, gent = crit(inp[0], inp[1], out)
        opt.zero_grad()
        loss.backward()
        sync.finish()
        opt.step()
        return loss

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ts, loss = [], None
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); loss = fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        # median over the steps: the eager legs launch ~1 000 kernels per step from Python and a single descheduled host
        # thread on a shared box shows up as a multi-ms outlier (the graph-replay leg below is the host-independent figure)
        t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(loss.detach())

    ms, loss = timed(step)
    n_params = sum(p.numel() for p in model.parameters())
    res = {"value": world * B * N / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "loss_rank0": loss, "n_gpus": world,
           "model": "%s, %d params" % (config_name, n_params), "batch_per_gpu": B, "points": N,
           "gradient_bytes_allreduced_per_step": 4 * n_params if world > 1 else 0,
           "includes": "H2D of the inputs, encoders, latent flows, priors, decoder, loss, backward, gradient all-reduce "
                       "(decoder arena overlapped with the rest of the backward), AMSGrad step; median of the timed steps, max over ranks"}
    if graphed:
        # the same step as two CUDA graphs (config key cuda_graph: forward+loss+backward | optimizer), _graphstep.py
        try:
            from dpf_nets_b200.lib.networks._graphstep import GraphedTrainStep
            gstep = GraphedTrainStep(model, crit, opt, allreduce=sync.finish, eager_steps=0)

            def run_graphed():
                return gstep(*[t.to(dev, non_blocking=True) for t in host])[0]
            gms, gloss = timed(run_graphed)
            res["cuda_graph"] = {"value": world * B * N / (gms * 1e-3), "unit": "points/s", "ms_per_step": gms, "loss_rank0": gloss}
        except Exception as e:      # reported, never fatal for the headline line
            res["cuda_graph"] = {"error": "%s: %s" % (type(e).__name__, e)}
    sync.remove()
    del model, opt
    torch.cuda.empty_cache()
    return res


def main():
    args = parse()
    # Libraries (NCCL's version banner, torchrun) write to fd 1; keep stdout to the ONE JSON line:
    # everything else goes to stderr, the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import builtins
    _print = builtins.print

    def json_print(*a, **k):
        k.setdefault("file", real_stdout)
        _print(*a, **k)
        real_stdout.flush()
    builtins.print = json_print
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.side_stream:      # A/B: everything on a non-default stream
            import torch as _t
            _t.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            with _t.cuda.stream(_t.cuda.Stream()):
                run_ours(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
