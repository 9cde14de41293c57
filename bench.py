#!/usr/bin/env python
"""bench.py -- TimesBlock forward windows/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload elec] [--impl native|reference]

A "step" is one forward of the TimesBlock stack (n_layers x (period search ->
fold -> Inception bank -> weighted aggregate + residual + shared LayerNorm)) over
one batch of synthetic pre-embedded windows that are already resident in HBM.
`value` = windows/s of that loop (whole job, all ranks).  `e2e` = windows/s of
the public API call a user makes (TimesNet.forward + NB-NLL) with HOST buffers:
each step copies x/y from pinned host memory, runs the forward and reads the
NLL back.  N > 1: one process per GPU (torchrun), weak scaling -- every rank
holds the full per-GPU batch, the only collective is the all-reduce of the
batch-summed amplitude spectrum inside the shared period search.

`--impl reference` times the reference algorithm's CPU path (the oracle port:
same ATen CPU kernels the reference's PyTorch code dispatches to; the reference
itself is not importable on the GPU box) on the same workload, all host threads.
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

# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernels, per launch, from the ncu --set full
# capture summarised under profiles/ (r1G); algorithmic bytes are in DESIGN.md section 4
NCU_TRAFFIC = {"elec": {"tc_conv4_kernel (block A, input computed once per window)": 4.5e6,
                        "tc_conv4_kernel (block B)": 21.0e6, "tc_mid_kernel": 37.9e6, "tc_tail_kernel": 61.7e6,
                        "tc_gemm2_kernel": 5.6e6, "spectrum_fft_kernel": 5.6e6, "unit": "bytes per launch",
                        "source": "profiles/r1G_ncu_full.csv"}}

METRIC = "timesblock_forward_windows_per_sec"
UNIT = "windows/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="elec", choices=["elec", "etth1", "traffic", "mid", "toy"])
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    return syn.WORKLOADS[args.workload]


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- #
# reference arm / cpu baseline: the oracle port on host cores
# --------------------------------------------------------------------------- #
def cpu_forward_factory(wl: syn.Workload, sample_B: int):
    sys.path.insert(0, str(ROOT / "oracle"))
    import flowtimes_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    w = model_state(wl, device="cpu")
    cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu", wl.min_period_threshold)
    x = syn.planted_series(sample_B, wl.T, wl.N, seed=0)
    y = syn.poisson_targets(sample_B, wl.H, wl.N, 5.0, seed=2)

    def step():
        with torch.no_grad():
            r, d = orc.timesnet_forward(x, w, cfg)
            return float(orc.nb_nll(y, r, d))
    return step


def model_state(wl: syn.Workload, device):
    """Seeded TimesNet state dict (keys = reference state_dict keys) for the workload."""
    shapes = dict(syn.stack_shapes(wl))
    C, N, L, H = wl.d_model, wl.N, wl.T, wl.H
    ctx = 32
    shapes.update({
        "embedding.gate": (1, 1, C), "embedding.value_embedding.weight": (C, N), "embedding.value_embedding.bias": (C,),
        "embedding.aux_norm.weight": (C,), "embedding.aux_norm.bias": (C,),
        "forecast_time_proj.weight": (H, L), "forecast_time_proj.bias": (H,),
        "mu_head.weight": (N, C), "mu_head.bias": (N,), "sigma_head.weight": (N, C), "sigma_head.bias": (N,),
        "series_embedding.weight": (N, ctx), "context_norm.weight": (ctx,), "context_norm.bias": (ctx,),
        "late_bias_norm.weight": (ctx,), "late_bias_norm.bias": (ctx,),
        "late_bias_head.weight": (H, ctx), "late_bias_head.bias": (H,), "late_bias_gate": (1, H, 1),
        "pre_embedding_norm.weight": (ctx + 1,), "pre_embedding_norm.bias": (ctx + 1,),
    })
    sd = syn.seeded_state(shapes, seed=0)
    return {k: v.to(device) for k, v in sd.items()}


def run_reference(args):
    """Reference arm: CPU path of the reference algorithm, full TimesNet.forward + NB-NLL."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload_of(args)
    sample_B = min(wl.B, 16)
    step = cpu_forward_factory(wl, sample_B)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample_B / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, **wl.as_dict(), **({"activation": args.activation} if args.activation != "gelu" else {}), "scope": "TimesNet.forward + NB-NLL on host cores",
                   "sample_windows_per_step": sample_B},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_B} of {wl.B} windows per step, full forward + NLL, fp32, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
def run_native(args):
    import torch.distributed as dist
    from timesnet_forecast import _native as nv
    from timesnet_forecast.losses import negative_binomial_nll
    from timesnet_forecast.models.timesnet import TimesNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    nv.load()
    wl = workload_of(args)
    sdt = syn.torch_dtype(wl.dtype)

    model = TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods,
                     kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0, activation=args.activation, mode=wl.mode,
                     d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio, min_period_threshold=wl.min_period_threshold,
                     use_checkpoint=False, stack_dtype=sdt)
    x_host = syn.planted_series(wl.B, wl.T, wl.N, seed=rank).pin_memory()
    y_host = syn.poisson_targets(wl.B, wl.H, wl.N, 5.0, seed=100 + rank).pin_memory()
    model(x_host[:1].to(dev))                                  # lazy build
    model.eval()
    model.load_state_dict(model_state(wl, dev), strict=True)
    model.check_finite = False                                 # keep the forward free of host syncs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------- resident scope: TimesBlock stack on pre-embedded features ----------
    n_rot = max(2, int(2 * 126e6 / (wl.B * wl.T * wl.d_model * (2 if sdt == torch.bfloat16 else 4))) + 1)
    n_rot = min(n_rot, 64)
    feats = [syn.planted_features(wl.B, wl.T, wl.d_model, seed=1000 * rank + i % 4).to(sdt).to(dev)
             for i in range(n_rot)]

    def stack_step(i):
        return model.stack_forward(feats[i % n_rot])

    for i in range(max(3, args.warmup)):
        stack_step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- pass A: eager launches with per-family CUDA-event timing inside the library ----
    nv.timing_enable(True)
    launches0 = nv.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        stack_step(i)
    ev1.record()
    barrier()
    eager_ms_total = ev0.elapsed_time(ev1)
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
    graphed = None
    if not args.no_graph:
        from timesnet_forecast.cuda_graphs import GraphedCallable
        graphed = GraphedCallable(model.stack_forward, [feats[0]])
        for i in range(max(3, args.warmup)):
            graphed(feats[i % n_rot])
        barrier()
        ev0.record()
        for i in range(args.steps):
            graphed(feats[i % n_rot])            # device copy of the step's input into the graph's buffer + replay
        ev1.record()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.finish()
    t = torch.tensor([ms_total, eager_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item() / args.steps
    eager_ms_step = t[1].item() / args.steps
    value = wl.B * world / (ms_step * 1e-3)
    group_periods = [list(b._last_plan.host().grp_period[: b._last_plan.host().n_groups]) for b in model.blocks]

    # ---------- e2e scope: public API with host buffers ----------
    e2e = None
    if not args.no_e2e:
        xd = torch.empty_like(x_host, device=dev)
        yd = torch.empty_like(y_host, device=dev)
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()

        def fwd_loss(xx, yy):
            rate, disp = model(xx)
            return negative_binomial_nll(yy, rate, disp)

        if args.no_graph:
            def e2e_step():
                xd.copy_(x_host, non_blocking=True)
                yd.copy_(y_host, non_blocking=True)
                loss_host.copy_(fwd_loss(xd, yd), non_blocking=True)
        else:
            # public streaming API: two graph slots, H2D of step i+1 on a copy stream overlaps the replay of step i;
            # every step still moves its own inputs host->device and its loss device->host
            from timesnet_forecast.cuda_graphs import PipelinedRunner
            runner = PipelinedRunner(fwd_loss, [xd, yd])
            loss_host = runner.results[0]

            def e2e_step():
                runner.submit(x_host, y_host)

        for _ in range(max(3, args.warmup)):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # the same host->device copies alone (nothing else on the GPU): the floor the host link puts under a step
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(5):
            xd.copy_(x_host, non_blocking=True)
            yd.copy_(y_host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_ms = c0.elapsed_time(c1) / 5
        h2d_bytes = int(x_host.numel() * 4 + y_host.numel() * 4)
        # ... and the same forward + loss alone (inputs already on the device)
        fwd_alone_ms = None
        if not args.no_graph:
            torch.cuda.synchronize()
            c0.record()
            for _ in range(5):
                runner._graphs[0].replay()
            c1.record()
            torch.cuda.synchronize()
            fwd_alone_ms = c0.elapsed_time(c1) / 5
        e2e = {"value": wl.B * world / (te.item() / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
               "ms_per_step": te.item() / args.steps,
               "h2d_alone_ms_per_step": h2d_ms, "h2d_alone_GBs": h2d_bytes / (h2d_ms * 1e-3) / 1e9,
               "forward_alone_ms_per_step": fwd_alone_ms,
               "scope": "TimesNet.forward + negative_binomial_nll" + ("" if args.no_graph else
                        " through PipelinedRunner (H2D of the next step overlaps the replay of the current one)"),
               "loss": float(loss_host)}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        # dominant kernel: the Inception conv chain (tensor-pipe bound by design)
        flops_fwd = sum(syn.stack_algorithmic_flops(syn.Workload(**{**wl.__dict__, "n_layers": 1}), gp)
                        for gp in group_periods)              # as-written conv FLOPs, SURVEY.md 8(d) K3 row
        flops_per_call = flops_fwd / max(1, len(group_periods))
        conv_avg_ms = conv_ms / max(1, conv_calls)
        achieved = flops_per_call / (conv_avg_ms * 1e-3) / 1e12 if conv_avg_ms > 0 else 0.0
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
        roofline = {"bound": "tensor", "kernel": "Inception chain of one TimesBlock (tc_gemm2, tc_conv4, tc_mid, tc_conv4, tc_tail)",
                    "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": None, "peak_source": peak_src + ", sustained bf16",
                    "flops_per_launch_group": flops_per_call, "avg_ms": conv_avg_ms, "calls": conv_calls,
                    "note": "as-written conv FLOPs of one TimesBlock (SURVEY 8d) / CUDA-event time of its Inception-chain "
                            "launches; proj o branch-out folding makes executed FLOPs 2.97x lower (elec), see profiles/",
                    "share_of_step": {"conv": conv_ms / eager_ms_total, "spectrum": spec_ms / eager_ms_total,
                                      "aggregate": agg_ms / eager_ms_total,
                                      "note": "shares of the eager pass (library CUDA events); aggregate is 0 when the "
                                              "fused tail (last 1x1 + aggregation + LayerNorm) runs inside the chain"}}
        # single kernels of the bf16 chain: executed (post weight-folding) FLOPs per call / measured time
        C_, F_, nb_ = wl.d_model, wl.ff, len(wl.kernel_set)
        mid_ = syn._mid(C_, F_, wl.bottleneck_ratio)
        taps_ = sum(kh * kw for kh, kw in wl.kernel_set)
        pos_per_call = wl.B * statistics.mean(sum(wl.T + ((-wl.T) % p) for p in gp) for gp in group_periods)
        mac_per_pos = {"s1_gemm": C_ * nb_ * mid_, "kk_a": taps_ * mid_ * mid_,
                       "mid": nb_ * mid_ * F_ + C_ * F_ + F_ * nb_ * mid_ + F_ * C_,
                       "kk_b": taps_ * mid_ * mid_, "s6_gemm": nb_ * mid_ * C_}
        kernels = []
        for name, (ms_f, calls) in chain.items():
            if calls:
                avg = ms_f / calls
                mac = mac_per_pos[name]
                # on the tc_conv4 route (mid = 32) the first 1x1 stage runs once per window, not once per period group
                units = wl.B * wl.T if (name == "s1_gemm" and mid_ == 32) else pos_per_call
                tf = 2.0 * mac * units / (avg * 1e-3) / 1e12
                label = "tail (last 1x1 + aggregate + LayerNorm)" if name == "s6_gemm" and not agg_calls else name
                kernels.append({"kernel": label, "avg_ms": avg, "executed_TFLOPs": tf, "frac_of_peak": tf / peak_tf,
                                "executed_mac_per_position": mac})
        roofline["chain_kernels"] = kernels
        roofline["executed_flops_per_launch_group"] = 2.0 * sum(mac_per_pos.values()) * pos_per_call
        # dram bytes of one tc_mid / tc_conv3 launch from the committed ncu --set full captures (profiles/)
        roofline["traffic"] = NCU_TRAFFIC.get(wl.name)
        e_bytes = 2 if sdt == torch.bfloat16 else 4
        hbm = []
        for name, ms_f, calls, per_call in (
                ("spectrum_fft" if search["fft"][1] else "period_search (all kernels)",
                 *(search["fft"] if search["fft"][1] else (spec_ms, spec_calls)),
                 wl.B * wl.T * wl.d_model * e_bytes + 4 * wl.B * (wl.T // 2 + 1) * wl.d_model),
                ("aggregate", agg_ms, agg_calls,
                 wl.B * wl.T * wl.d_model * e_bytes * (statistics.mean(len(g) for g in group_periods) + 2))):
            if calls:
                gbs = per_call / (ms_f / calls * 1e-3) / 1e9
                hbm.append({"kernel": name, "achieved_GBs": gbs, "frac": gbs / float(peaks["hbm_gbs"]),
                            "avg_ms": ms_f / calls, "algorithmic_bytes": per_call})
        roofline["hbm_kernels"] = hbm
        roofline["search_kernels"] = [{"kernel": n, "avg_ms": ms_f / calls} for n, (ms_f, calls) in search.items() if calls]
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            sample_B = min(wl.B, 8)
            step = cpu_forward_factory(wl, sample_B)
            step()
            t0 = time.perf_counter()
            n = 0
            while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
                step()
                n += 1
            dt = (time.perf_counter() - t0) / n
            cores = torch.get_num_threads()
            cpu_baseline = {"value": sample_B / dt, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n} forwards of {sample_B}/{wl.B} windows, TimesNet.forward + NLL, fp32, "
                                      f"{cores} threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": wl.name, **wl.as_dict(), "global_batch": wl.B * world, "parallelism": f"dp{world}",
                       "scope": "TimesBlock stack (n_layers x (TimesBlock + shared LayerNorm)), features resident"
                                + ("" if args.no_graph else "; step = copy into the graph input + CUDA-graph replay"),
                       "l2": f"inputs rotate over {n_rot} buffers (> 2x L2); fp32 intermediates >> L2",
                       "selected_periods": group_periods},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "launch_mode": "eager" if args.no_graph else "cuda_graph",
            "eager_ms_per_step": eager_ms_step, "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
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
