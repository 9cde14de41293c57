#!/usr/bin/env python
"""bench.py -- TimesBlock forward windows/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload elec|etth1|traffic|recursive]
                    [--dtype f32|bf16] [--impl native|reference]

Direct workloads (elec = default, etth1, traffic): a "step" is one forward of the TimesBlock stack
(n_layers x (period search -> fold -> Inception bank -> weighted aggregate + residual + shared
LayerNorm)) over one batch of synthetic pre-embedded windows already resident in HBM.  `value` =
windows/s of that loop (whole job, all ranks).  `e2e` = windows/s of the public API call a user makes
(TimesNet.forward + NB-NLL) with HOST buffers: each step copies x / y from pinned host memory, runs the
forward and reads the NLL back; `e2e_forecast` is the serving variant that reads rate + dispersion back.

Recursive workload (BASELINE config 5): a "step" is one 28-step rolling forecast (predict.py:307-342) over
the rank's series; `value` = window-steps/s (series x 28 forwards per second), `series_per_sec` rides along.

N > 1: one process per GPU (torchrun), weak scaling -- every rank holds the full per-GPU batch; the only
collective is the all-reduce of the batch-summed amplitude spectrum inside the shared period search (opt-in,
`parallel.share_period_search`).  The N > 1 line also carries a `strong_scaling` block (global batch of the
1-GPU configuration sharded over the ranks) without changing the default metric.

`--impl reference` times the reference algorithm's CPU path (the oracle port: same ATen CPU kernels the
reference's PyTorch code dispatches to; the reference itself is not importable on the GPU box) on the same
workload, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))

import torch  # noqa: E402

import flowtimes_synth as syn  # noqa: E402

METRIC = "timesblock_forward_windows_per_sec"
UNIT = "windows/s"
ID_EMBED = 32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="elec", choices=["elec", "etth1", "traffic", "recursive", "mid", "toy"])
    ap.add_argument("--dtype", default=None, choices=["f32", "bf16"], help="activation dtype of the TimesBlock stack "
                    "(default: the workload's; recursive defaults to bf16, BASELINE names none)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (0 = the workload's)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block of the N > 1 line")
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: how the spectrum sums are exchanged (NVLink peer mailbox inside the selection kernel, or NCCL)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager launch path instead of CUDA-graph replay")
    ap.add_argument("--activation", default="gelu", choices=["gelu", "relu"],
                    help="diagnostic only: relu removes the GELU cost from the epilogues (the metric is quoted on gelu)")
    return ap.parse_args()


# --------------------------------------------------------------------------- #
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 10 ms through NVML (the timed regions are tens of milliseconds long)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, v in names.items():
                    if bits & v:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(0.01)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def workload_of(args) -> syn.Workload:
    wl = syn.WORKLOADS[args.workload]
    over = {}
    if args.dtype:
        over["dtype"] = args.dtype
    elif wl.name == "recursive":
        over["dtype"] = "bf16"
    if args.batch > 0:
        over["B"] = args.batch
    elif wl.name == "recursive" and args.gpus > 1:
        over["B"] = wl.B // args.gpus            # 30 000 series over the ranks (3 750 per GPU at 8)
    return syn.Workload(**{**wl.__dict__, **over}) if over else wl


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def committed_ncu_summary(workload: str):
    """Per-kernel numbers of the latest committed `ncu --set full` capture for this workload
    (profiles/*_ncu_summary.json, written by profiles/summarize.py; NOT measured by this run)."""
    best = None
    for p in sorted((ROOT / "profiles").glob("*_ncu_summary.json")):
        try:
            d = json.loads(p.read_text())
        except Exception:
            continue
        if d.get("workload") == workload:
            best = (p, d)
    if best is None:
        return None
    p, d = best
    d = dict(d)
    d["source"] = f"profiles/{p.name} (committed ncu --set full capture; not measured by this run)"
    return d


# --------------------------------------------------------------------------- #
# seeded model state shared by both arms
# --------------------------------------------------------------------------- #
def model_state(wl: syn.Workload, device):
    """Seeded TimesNet state dict (keys = reference state_dict keys) for the workload."""
    shapes = dict(syn.stack_shapes(wl))
    C, N, L, H = wl.d_model, wl.N, wl.T, wl.H
    steps = H if wl.mode == "direct" else 1
    static_out = wl.static_features                     # static_proj_dim None -> proj dim = input features
    ctx = ID_EMBED + static_out
    shapes.update({
        "embedding.gate": (1, 1, C), "embedding.value_embedding.weight": (C, N), "embedding.value_embedding.bias": (C,),
        "embedding.aux_norm.weight": (C,), "embedding.aux_norm.bias": (C,),
        "forecast_time_proj.weight": (H, L), "forecast_time_proj.bias": (H,),
        "mu_head.weight": (N, C), "mu_head.bias": (N,), "sigma_head.weight": (N, C), "sigma_head.bias": (N,),
        "series_embedding.weight": (N, ID_EMBED), "context_norm.weight": (ctx,), "context_norm.bias": (ctx,),
        "late_bias_norm.weight": (ctx,), "late_bias_norm.bias": (ctx,),
        "late_bias_head.weight": (steps, ctx), "late_bias_head.bias": (steps,), "late_bias_gate": (1, steps, 1),
        "pre_embedding_norm.weight": (ctx + 1,), "pre_embedding_norm.bias": (ctx + 1,),
    })
    if static_out > 0:
        shapes.update({"static_proj.weight": (static_out, wl.static_features), "static_proj.bias": (static_out,),
                       "static_norm.weight": (static_out,), "static_norm.bias": (static_out,)})
    if wl.context_rank > 0:
        shapes.update({"context_coeff.weight": (wl.context_rank, ctx), "context_coeff.bias": (wl.context_rank,),
                       "temporal_context.scale": ()})
    sd = syn.seeded_state(shapes, seed=0)
    return {k: v.to(device) for k, v in sd.items()}


def recursive_inputs(wl: syn.Workload, B: int, seed: int):
    """Config 5 inputs: counts ~ Poisson(4), per-series statics ~ N(0, 1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.poisson(torch.full((B, wl.T, wl.N), 4.0), generator=g)
    static = torch.randn(B, wl.N, wl.static_features, generator=g) if wl.static_features > 0 else None
    return x, static


# --------------------------------------------------------------------------- #
# reference arm / cpu baseline: the oracle port on host cores
# --------------------------------------------------------------------------- #
def cpu_step_factory(wl: syn.Workload, sample_B: int, rec_steps: int = 2):
    """One bounded sample of the workload on the host: returns (step_fn, units_per_step, description)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import flowtimes_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    w = model_state(wl, device="cpu")
    cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu", wl.min_period_threshold,
                       1e-3, wl.context_rank > 0, wl.context_rank)
    if wl.mode == "recursive":
        x, static = recursive_inputs(wl, sample_B, seed=0)

        def step():
            with torch.no_grad():
                r, d = orc.forecast_recursive(x, rec_steps, w, cfg, series_static=static,
                                              series_ids=torch.arange(wl.N))
                return float(r.sum())
        return step, sample_B * rec_steps, (f"{rec_steps} of {wl.H} rolling steps over {sample_B} of {wl.B} series, "
                                            "full forward per step, fp32")
    x = syn.planted_series(sample_B, wl.T, wl.N, seed=0)
    y = syn.poisson_targets(sample_B, wl.H, wl.N, 5.0, seed=2)

    def step():
        with torch.no_grad():
            r, d = orc.timesnet_forward(x, w, cfg)
            return float(orc.nb_nll(y, r, d))
    return step, sample_B, f"{sample_B} of {wl.B} windows per step, full forward + NLL, fp32"


def run_reference(args):
    """Reference arm: CPU path of the reference algorithm, full TimesNet.forward (+ NB-NLL) per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload_of(args)
    sample_B = min(wl.B, 256 if wl.mode == "recursive" else (4 if wl.name == "traffic" else 16))
    step, units, desc = cpu_step_factory(wl, sample_B)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = units / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, **wl.as_dict(), **({"activation": args.activation} if args.activation != "gelu" else {}),
                   "scope": "TimesNet.forward + NB-NLL on host cores" if wl.mode == "direct"
                            else "rolling one-step forecast (TimesNet.forward per step) on host cores",
                   "sample_units_per_step": units,
                   "sampling_note": "the shared period search sees the sampled batch, not the full one (SURVEY 9.11)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{desc}, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
def timed_loop(fn, steps, barrier):
    """K calls of fn(i) between two CUDA events on the current stream, barrier + synchronize on both sides."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def run_native(args):
    import torch.distributed as dist
    from timesnet_forecast import _native as nv
    from timesnet_forecast.losses import negative_binomial_nll
    from timesnet_forecast.models.timesnet import TimesNet
    from timesnet_forecast.parallel import share_period_search

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stdout_fd = None
    if world > 1:
        # the contract is ONE JSON line on stdout: NCCL prints its version banner there (NCCL_DEBUG=VERSION on the pool's
        # multi-GPU boxes), so stdout points at stderr until the line is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    nv.load()
    args.gpus = world
    wl = workload_of(args)
    sdt = syn.torch_dtype(wl.dtype)
    recursive = wl.mode == "recursive"

    model = TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods,
                     kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0, activation=args.activation, mode=wl.mode,
                     d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio, min_period_threshold=wl.min_period_threshold,
                     use_checkpoint=False, stack_dtype=sdt, use_zero_mean_context=wl.context_rank > 0,
                     context_rank=wl.context_rank, context_scale=0.05)
    if world > 1:
        share_period_search(model, transport=args.transport)    # the shared search of SURVEY 8e is opt-in
    ids = torch.arange(wl.N, device=dev)
    if recursive:
        x_cpu, st_cpu = recursive_inputs(wl, wl.B, seed=rank)
        x_host = x_cpu.pin_memory()
        static_host = st_cpu.pin_memory() if st_cpu is not None else None
        y_host = None
        static_dev = static_host.to(dev) if static_host is not None else None
        model.eval()
        model(x_host[:1].to(dev), series_static=None if static_dev is None else static_dev[:1], series_ids=ids)
    else:
        x_host = syn.planted_series(wl.B, wl.T, wl.N, seed=rank).pin_memory()
        y_host = syn.poisson_targets(wl.B, wl.H, wl.N, 5.0, seed=100 + rank).pin_memory()
        static_host = static_dev = None
        model.eval()
        model(x_host[:1].to(dev))                               # lazy build
    model.load_state_dict(model_state(wl, dev), strict=True)
    model.check_finite = False                                  # keep the forward free of host syncs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------- resident scope ----------
    e_bytes = 2 if sdt == torch.bfloat16 else 4
    sampler = ClockSampler(local)
    if recursive:
        from timesnet_forecast.predict import RecursiveForecaster
        x_dev = x_host.to(dev)
        eager_runner = RecursiveForecaster(model, x_dev, wl.H, series_static=static_dev, series_ids=ids, graph=False)
        graph_runner = None if args.no_graph else RecursiveForecaster(model, x_dev, wl.H, series_static=static_dev,
                                                                      series_ids=ids, graph=True)

        def eager_step(i):
            eager_runner.run(x_dev)

        def graph_step(i):
            graph_runner.run(x_dev)
        n_rot = 1
        l2_note = "every rolling step rewrites > 2x L2 of intermediates (fp32/bf16 chain buffers >> 126 MB)"
        units_per_step = wl.B * wl.H
    else:
        n_rot = max(2, int(2 * 126e6 / (wl.B * wl.T * wl.d_model * e_bytes)) + 1)
        n_rot = min(n_rot, 64)
        feats = [syn.planted_features(wl.B, wl.T, wl.d_model, seed=1000 * rank + i % 4).to(sdt).to(dev)
                 for i in range(n_rot)]

        def eager_step(i):
            return model.stack_forward(feats[i % n_rot])
        graphed = None
        l2_note = f"inputs rotate over {n_rot} buffers (> 2x L2); intermediates >> L2"
        units_per_step = wl.B

    for i in range(max(3, args.warmup)):
        eager_step(i)
    barrier()
    sampler.start()
    # ---- pass A: eager launches with per-family CUDA-event timing inside the library ----
    nv.timing_enable(True)
    launches0 = nv.launch_count()
    eager_ms_total = timed_loop(eager_step, args.steps, barrier)
    launches = nv.launch_count() - launches0
    conv_ms, conv_calls = nv.timing_read(nv.FAM_CONV)
    spec_ms, spec_calls = nv.timing_read(nv.FAM_SPECTRUM)
    agg_ms, agg_calls = nv.timing_read(nv.FAM_AGGREGATE)
    chain = {name: nv.timing_read(fam) for name, fam in (
        ("s1_gemm", nv.FAM_S1), ("kk_a", nv.FAM_KK_A), ("mid", nv.FAM_MID), ("kk_b", nv.FAM_KK_B), ("s6_gemm", nv.FAM_S6))}
    search = {name: nv.timing_read(fam) for name, fam in (
        ("fft", nv.FAM_FFT), ("median", nv.FAM_MEDIAN), ("select", nv.FAM_SELECT))}
    nv.timing_enable(False)
    ms_total = eager_ms_total
    # ---- pass B: the same step replayed from a CUDA graph (no host round trip exists on the path) ----
    if not args.no_graph:
        if recursive:
            step_fn = graph_step
        else:
            from timesnet_forecast.cuda_graphs import GraphedCallable
            graphed = GraphedCallable(model.stack_forward, [feats[0]], params_of=model)

            def step_fn(i):
                graphed(feats[i % n_rot])        # device copy of the step's input into the graph's buffer + replay
        for i in range(max(3, args.warmup)):
            step_fn(i)
        ms_total = timed_loop(step_fn, args.steps, barrier)
    clocks = sampler.finish()
    t = torch.tensor([ms_total, eager_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item() / args.steps
    eager_ms_step = t[1].item() / args.steps
    value = units_per_step * world / (ms_step * 1e-3)
    plans = [b._last_plan.host() for b in model.blocks]
    group_periods = [list(h.grp_period[: h.n_groups]) for h in plans]
    same_periods = None
    if world > 1:                                               # the shared search must agree on every rank
        flat = torch.tensor([p for gp in group_periods for p in gp + [-1]], dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(flat) for _ in range(world)]
        try:
            dist.all_gather(gathered, flat)
            same_periods = all(torch.equal(g, gathered[0]) for g in gathered)
        except Exception:
            same_periods = False
        assert same_periods, "ranks selected different periods"

    # ---------- strong scaling block (N > 1): the 1-GPU global batch sharded over the ranks ----------
    strong = None
    if world > 1 and not recursive and not args.no_strong and wl.B % world == 0:
        Bs = wl.B // world
        sfeats = [f[:Bs].contiguous() for f in feats[: min(n_rot, 8)]]
        if args.no_graph:
            def strong_step(i):
                model.stack_forward(sfeats[i % len(sfeats)])
        else:
            from timesnet_forecast.cuda_graphs import GraphedCallable
            sg = GraphedCallable(model.stack_forward, [sfeats[0]], params_of=model)

            def strong_step(i):
                sg(sfeats[i % len(sfeats)])
        for i in range(max(3, args.warmup)):
            strong_step(i)
        ts = torch.tensor([timed_loop(strong_step, args.steps, barrier)], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sms = ts.item() / args.steps
        strong = {"global_batch": wl.B, "per_gpu_batch": Bs, "ms_per_step": sms, "value": wl.B / (sms * 1e-3), "unit": UNIT,
                  "note": "same stack step with the 1-GPU global batch sharded over the ranks (SURVEY 8d); the default "
                          "metric above stays weak scaling"}

    # ---------- e2e scope: public API with host buffers ----------
    e2e = None
    e2e_forecast = None
    if not args.no_e2e and recursive:
        rates_host = torch.empty(wl.B, wl.H, wl.N).pin_memory()
        disps_host = torch.empty(wl.B, wl.H, wl.N).pin_memory()
        runner = graph_runner if graph_runner is not None else eager_runner
        xd = torch.empty_like(x_host, device=dev)

        def e2e_step(i):
            xd.copy_(x_host, non_blocking=True)
            if static_host is not None:
                static_dev.copy_(static_host, non_blocking=True)
            r, d = runner.run(xd)
            rates_host.copy_(r, non_blocking=True)
            disps_host.copy_(d, non_blocking=True)
        for i in range(3):
            e2e_step(i)
        te = torch.tensor([timed_loop(e2e_step, max(2, args.steps // 2), barrier)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ems = te.item() / max(2, args.steps // 2)
        h2d = int(x_host.numel() * 4 + (static_host.numel() * 4 if static_host is not None else 0))
        e2e = {"value": units_per_step * world / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(2 * rates_host.numel() * 4), "ms_per_step": ems,
               "series_per_sec": wl.B * world / (ems * 1e-3),
               "scope": "RecursiveForecaster.run: last_seq + statics from pinned host memory, 28 rolling forwards "
                        "(one captured graph replayed), rate + dispersion [B, H, N] read back"}
    elif not args.no_e2e:
        xd = torch.empty_like(x_host, device=dev)
        yd = torch.empty_like(y_host, device=dev)
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()

        def fwd_loss(xx, yy):
            rate, disp = model(xx)
            return negative_binomial_nll(yy, rate, disp)

        if args.no_graph:
            def e2e_step(i):
                xd.copy_(x_host, non_blocking=True)
                yd.copy_(y_host, non_blocking=True)
                loss_host.copy_(fwd_loss(xd, yd), non_blocking=True)
        else:
            # public streaming API: two graph slots, H2D of step i+1 on a copy stream overlaps the replay of step i;
            # every step still moves its own inputs host->device and its loss device->host
            from timesnet_forecast.cuda_graphs import PipelinedRunner
            runner = PipelinedRunner(fwd_loss, [xd, yd], params_of=model)
            loss_host = runner.results[0]

            def e2e_step(i):
                runner.submit(x_host, y_host)

        for i in range(max(3, args.warmup)):
            e2e_step(i)
        te = torch.tensor([timed_loop(e2e_step, args.steps, barrier)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # the same host->device copies alone (nothing else on the GPU): the floor the host link puts under a step
        def copies(i):
            xd.copy_(x_host, non_blocking=True)
            yd.copy_(y_host, non_blocking=True)
        h2d_ms = timed_loop(copies, 5, barrier) / 5
        h2d_bytes = int(x_host.numel() * 4 + y_host.numel() * 4)
        th = torch.tensor([h2d_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)          # all ranks copy at the same time: the shared host link
        # ... and the same forward + loss alone (inputs already on the device)
        fwd_alone_ms = None
        if not args.no_graph:
            fwd_alone_ms = timed_loop(lambda i: runner._graphs[0].replay(), 5, barrier) / 5
        e2e = {"value": wl.B * world / (te.item() / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
               "ms_per_step": te.item() / args.steps,
               "h2d_alone_ms_per_step": h2d_ms, "h2d_alone_GBs": h2d_bytes / (h2d_ms * 1e-3) / 1e9,
               "h2d_all_ranks_ms_per_step": th.item(),
               "h2d_aggregate_GBs": h2d_bytes * world / (th.item() * 1e-3) / 1e9,
               "host_link_bound_windows_per_sec": wl.B * world / (th.item() * 1e-3),
               "forward_alone_ms_per_step": fwd_alone_ms,
               "scope": "TimesNet.forward + negative_binomial_nll" + ("" if args.no_graph else
                        " through PipelinedRunner (H2D of the next step overlaps the replay of the current one)"),
               "loss": float(loss_host)}
        # serving variant: the caller reads rate + dispersion back instead of a loss
        if not args.no_graph:
            def fwd_pack(xx):
                rate, disp = model(xx)
                return torch.stack([rate, disp])
            frun = PipelinedRunner(fwd_pack, [xd], params_of=model)
            for i in range(3):
                frun.submit(x_host)
            tf = torch.tensor([timed_loop(lambda i: frun.submit(x_host), args.steps, barrier)], dtype=torch.float64,
                              device=dev)
            if world > 1:
                dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            e2e_forecast = {"value": wl.B * world / (tf.item() / args.steps * 1e-3), "unit": UNIT,
                            "ms_per_step": tf.item() / args.steps, "h2d_bytes_per_step": int(x_host.numel() * 4),
                            "d2h_bytes_per_step": int(2 * wl.B * wl.H * wl.N * 4),
                            "scope": "TimesNet.forward, rate + dispersion [B, H, N] read back to pinned host memory"}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        # dominant kernel family: the Inception conv chain (tensor-pipe bound by design)
        one_layer = syn.Workload(**{**wl.__dict__, "n_layers": 1})
        flops_fwd = sum(syn.stack_algorithmic_flops(one_layer, gp) for gp in group_periods)   # as-written, SURVEY 8(d) K3
        flops_per_call = flops_fwd / max(1, len(group_periods))
        conv_avg_ms = conv_ms / max(1, conv_calls)
        achieved = flops_per_call / (conv_avg_ms * 1e-3) / 1e12 if conv_avg_ms > 0 else 0.0
        peak_burst = float(peaks.get("bf16_tflops", 1590.0))
        peak_sust = float(peaks.get("bf16_tflops_sustained", peak_burst))
        C_, F_, nb_ = wl.d_model, wl.ff, len(wl.kernel_set)
        mid_ = syn._mid(C_, F_, wl.bottleneck_ratio)
        taps_ = sum(kh * kw for kh, kw in wl.kernel_set)
        pos_per_call = wl.B * statistics.mean(sum(wl.T + ((-wl.T) % p) for p in gp) for gp in group_periods)
        mac_per_pos = {"s1_gemm": C_ * nb_ * mid_, "kk_a": taps_ * mid_ * mid_,
                       "mid": nb_ * mid_ * F_ + C_ * F_ + F_ * nb_ * mid_ + F_ * C_,
                       "kk_b": taps_ * mid_ * mid_, "s6_gemm": nb_ * mid_ * C_}
        exec_flops = 2.0 * sum(mac_per_pos.values()) * pos_per_call
        achieved_exec = exec_flops / (conv_avg_ms * 1e-3) / 1e12 if conv_avg_ms > 0 else 0.0
        ncu = committed_ncu_summary(wl.name)
        roofline = {"bound": "tensor",
                    "kernel": "Inception chain of one TimesBlock (all launches between the period search and the block output)",
                    "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s", "frac": achieved / peak_burst,
                    "peak_source": peak_src + ", burst bf16 (the timed region is tens of ms at max clock)",
                    "frac_vs_sustained_peak": achieved / peak_sust, "peak_sustained": peak_sust,
                    "achieved_executed": achieved_exec, "frac_executed": achieved_exec / peak_burst,
                    "frac_executed_vs_sustained_peak": achieved_exec / peak_sust,
                    "traffic": None if ncu is None else ncu.get("dram_bytes_per_launch"),
                    "traffic_source": None if ncu is None else ncu["source"],
                    "flops_per_launch_group": flops_per_call, "executed_flops_per_launch_group": exec_flops,
                    "avg_ms": conv_avg_ms, "calls": conv_calls,
                    "note": "achieved / frac = AS-WRITTEN conv FLOPs of one TimesBlock (SURVEY 8d) over the CUDA-event time of "
                            "its Inception-chain launches; proj o branch-out folding makes the EXECUTED FLOPs lower "
                            "(achieved_executed / frac_executed); with an fp32 stack on the tensor cores every product is "
                            "6 bf16 MMAs (3-way split), not counted here; hardware pipe activity per kernel is under "
                            "`ncu` (committed capture)",
                    "share_of_step": {"conv": conv_ms / eager_ms_total, "spectrum": spec_ms / eager_ms_total,
                                      "aggregate": agg_ms / eager_ms_total,
                                      "note": "shares of the eager pass (library CUDA events); aggregate is 0 when the "
                                              "fused tail (last 1x1 + aggregation + LayerNorm) runs inside the chain"}}
        if ncu is not None:
            roofline["ncu"] = {k: ncu[k] for k in ("kernels", "step_dram_bytes", "compulsory_bytes", "dram_over_compulsory")
                               if k in ncu}
            roofline["ncu"]["source"] = ncu["source"]
        kernels = []
        for name, (ms_f, calls) in chain.items():
            if calls:
                avg = ms_f / calls
                mac = mac_per_pos[name]
                # on the tc_conv4 route (mid = 32, bf16) the first 1x1 stage runs once per window, not once per period group
                units = wl.B * wl.T if (name == "s1_gemm" and mid_ == 32 and sdt == torch.bfloat16) else pos_per_call
                tf = 2.0 * mac * units / (avg * 1e-3) / 1e12
                label = "tail (last 1x1 + aggregate + LayerNorm)" if name == "s6_gemm" and not agg_calls else name
                kernels.append({"kernel": label, "avg_ms": avg, "executed_TFLOPs": tf, "frac_of_burst_peak": tf / peak_burst,
                                "executed_mac_per_position": mac})
        roofline["chain_kernels"] = kernels
        hbm = []
        Fq = wl.T // 2 + 1
        one_launch = (not recursive) and nv.dft_basis(feats[0]) is not None      # tensor-core search (csrc/tc_dft.cu)
        for name, ms_f, calls, per_call in (
                ("period_search = tc_dft_kernel (spectrum GEMM + channel median + selection tail, ONE launch)" if one_launch
                 else "spectrum_fft" if search["fft"][1] else "period_search (all kernels)",
                 *(search["fft"] if search["fft"][1] else (spec_ms, spec_calls)),
                 wl.B * wl.T * wl.d_model * e_bytes + 4 * wl.B * Fq),        # read x once + medians out
                ("aggregate", agg_ms, agg_calls,
                 wl.B * wl.T * wl.d_model * e_bytes * (statistics.mean(len(g) for g in group_periods) + 2))):
            if calls:
                gbs = per_call / (ms_f / calls * 1e-3) / 1e9
                hbm.append({"kernel": name, "achieved_GBs": gbs, "frac": gbs / float(peaks["hbm_gbs"]),
                            "avg_ms": ms_f / calls, "algorithmic_bytes": per_call})
        if one_launch and hbm:
            # SURVEY 8d classes K1 as HBM-bound (read x once); the one-launch search is not: its spectra take ~12 us
            # (a tcgen05 GEMM against the three-plane DFT basis, L2-resident, plus a register-sort median) and the
            # one-CTA selection tail another ~13 us of pure latency.  The GEMM's executed FLOPs are reported beside it.
            mt = (Fq + 63) // 64
            dft_flops = 2.0 * 3 * (mt * 128) * wl.T * wl.d_model * wl.B
            hbm[0]["executed_TFLOPs"] = dft_flops / (hbm[0]["avg_ms"] * 1e-3) / 1e12
            hbm[0]["note"] = ("latency-bound, not HBM-bound: eager CUDA-event time of the whole search (spectra + tail); "
                              "profiles/ holds the in-graph timeline and the tail's phase trace")
        roofline["hbm_kernels"] = hbm
        roofline["search_kernels"] = [{"kernel": n, "avg_ms": ms_f / calls} for n, (ms_f, calls) in search.items() if calls]
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            sample_B = min(wl.B, 128 if recursive else (2 if wl.name == "traffic" else 8))
            step, units, desc = cpu_step_factory(wl, sample_B)
            step()
            t0 = time.perf_counter()
            n = 0
            while n < 2 or (time.perf_counter() - t0 < 10.0 and n < 50):
                step()
                n += 1
            dt = (time.perf_counter() - t0) / n
            cores = torch.get_num_threads()
            cpu_baseline = {"value": units / dt, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n} runs of: {desc}, {cores} threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if (recursive and world > 1 and args.batch <= 0) else "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": wl.name, **wl.as_dict(), "global_batch": wl.B * world, "parallelism": f"dp{world}",
                       "scope": ("rolling one-step forecast: 28 x (TimesNet.forward + device-side append/roll), series "
                                 "resident; value counts window-steps (series x 28)" if recursive else
                                 "TimesBlock stack (n_layers x (TimesBlock + shared LayerNorm)), features resident")
                                + ("" if args.no_graph else "; step = CUDA-graph replay"),
                       "l2": l2_note, "selected_periods": group_periods,
                       "spectrum_exchange": None if world == 1 else (
                           "NVLink peer mailbox inside the selection kernel" if model.period_selector.peer_comm is not None
                           else "NCCL all-reduce")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "launch_mode": "eager" if args.no_graph else "cuda_graph",
            "eager_ms_per_step": eager_ms_step, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        if recursive:
            line["series_per_sec"] = wl.B * world / (ms_step * 1e-3)
        if e2e_forecast is not None:
            line["e2e_forecast"] = e2e_forecast
        if strong is not None:
            line["strong_scaling"] = strong
        if same_periods is not None:
            line["periods_identical_on_all_ranks"] = same_periods
        if stdout_fd is not None:
            sys.stdout.flush()
            os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: communicator destruction with captured graphs that still hold
        # NCCL nodes can block forever (seen on the 2-GPU box).  Everything is flushed and synchronised here.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
