"""Pin the oracle against the real reference and write the golden fixtures.

Runs ONLY in the build container (needs ``/root/reference/src``).  For every
case it (1) runs the unmodified reference PyTorch modules on CPU fp32/bf16,
(2) runs ``oracle/flowtimes_oracle.py`` on the same seeded inputs and weights,
(3) asserts agreement (bit-exact for integer outputs, <= 1e-6 abs for floats --
both are the same ATen kernels, usually 0.0), and (4) stores the REFERENCE's
outputs under ``tests/golden/`` so the GPU box (no ``/root/reference``) can
check both the oracle and the CUDA path against them.

    python oracle/make_golden.py            # rewrites tests/golden/*.pt
"""
from __future__ import annotations

import math
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
REF_SRC = Path("/root/reference/src")
if not REF_SRC.exists():
    raise SystemExit("make_golden.py needs /root/reference/src (build container only)")
sys.path.insert(0, str(REF_SRC))

import flowtimes_oracle as orc                      # noqa: E402
import flowtimes_synth as syn                       # noqa: E402
from timesnet_forecast.models import timesnet as ref  # noqa: E402  (the real reference)
from timesnet_forecast.losses import negative_binomial_nll as ref_nll  # noqa: E402
from timesnet_forecast.predict import forecast_recursive_batch as ref_recursive  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
torch.set_num_threads(max(1, os.cpu_count() or 1))


class FixedSelector(torch.nn.Module):
    """Same fake the reference tests inject (tests/test_times_block.py:14-30)."""

    def __init__(self, periods, amplitudes, pmax=None, min_period_threshold=None):
        super().__init__()
        self.p = torch.as_tensor(periods, dtype=torch.long)
        self.a = torch.as_tensor(amplitudes, dtype=torch.float32)
        if pmax is not None:
            self.pmax = pmax
        if min_period_threshold is not None:
            self.min_period_threshold = min_period_threshold

    def forward(self, x):
        a = self.a.to(x.dtype)
        a = a.unsqueeze(0) if a.dim() == 1 else a
        a = a.expand(x.size(0), -1) if a.size(0) == 1 and x.size(0) > 1 else a
        return self.p, a


def check(name, got, want, tol=1e-5):
    """Integer outputs bit-exact; floats within ``tol`` relative to max|want|.

    The reference runs its convs on channels_last weights (oneDNN picks other
    kernels than for the oracle's contiguous tensors), so conv outputs differ by
    ~1e-6 relative even though the arithmetic is the same.
    """
    if want.dtype in (torch.long, torch.int32, torch.bool):
        assert torch.equal(got, want), f"{name}: integer mismatch {got} vs {want}"
        return 0.0
    if want.numel() == 0:
        assert got.numel() == 0, name
        return 0.0
    nan_w, nan_g = torch.isnan(want.float()), torch.isnan(got.float())
    assert torch.equal(nan_w, nan_g), f"{name}: NaN pattern differs"
    if bool(nan_w.all()):
        return 0.0
    gotf = torch.where(nan_w, torch.zeros_like(got.float()), got.float())
    wantf = torch.where(nan_w, torch.zeros_like(want.float()), want.float())
    scale = max(1.0, wantf.abs().max().item())
    d = (gotf - wantf).abs().max().item() / scale
    assert d <= tol, f"{name}: oracle differs from reference by {d} (relative to max)"
    return d


def subsample(t: torch.Tensor, n: int = 4096) -> torch.Tensor:
    flat = t.reshape(-1).float()
    step = max(1, flat.numel() // n)
    return flat[::step].clone()


# --------------------------------------------------------------------------- #
def selector_cases():
    cases = {}
    # the reference's own known-answer tests (tests/test_fft_period_selector.py)
    torch.manual_seed(0)
    L = 256
    t = torch.arange(L, dtype=torch.float32)
    xs = []
    for _ in range(2):
        cols = []
        for _ in range(3):
            s = 3.0 * torch.sin(2 * math.pi * 4 * t / L) + 1.5 * torch.sin(2 * math.pi * 8 * t / L)
            cols.append(s + 0.01 * torch.randn_like(t))
        xs.append(torch.stack(cols, dim=1))
    cases["ref_shared_L256"] = (torch.stack(xs, 0), 2, L, 1, [64, 32])
    L = 64
    t = torch.arange(L, dtype=torch.float32)
    s = 2.0 * torch.sin(2 * math.pi * 2 * t / L) + 1.0 * torch.sin(2 * math.pi * 20 * t / L)
    cases["ref_bounds_L64"] = (s.view(1, L, 1), 2, 16, 5, [16, 5])
    L = 32
    t = torch.arange(L, dtype=torch.float32)
    cases["ref_zero_k"] = (torch.sin(2 * math.pi * 3 * t / L).view(1, L, 1), 0, L, 1, [])
    L = 28
    t = torch.arange(L, dtype=torch.float32)
    cases["ref_min_cycles_L28"] = (torch.sin(2 * math.pi * t / 7).view(1, L, 1), 3, L, 1, None)
    g = torch.Generator().manual_seed(7)
    cases["odd_L150_C4"] = (torch.randn(2, 150, 4, generator=g), 3, 150, 1, None)
    cases["prime_L37_C6"] = (torch.randn(3, 37, 6, generator=g), 4, 37, 2, None)
    cases["even_C_lower_median"] = (torch.randn(2, 40, 2, generator=g), 3, 40, 1, None)
    out = {}
    for name, (x, k, pmax, mpt, expect) in cases.items():
        sel = ref.FFTPeriodSelector(k_periods=k, pmax=pmax, min_period_threshold=mpt)
        p, a = sel(x)
        o = orc.select_periods(x, k, pmax, mpt)
        check(name + ".periods", o.periods, p)
        check(name + ".amps", o.amplitudes, a, 0.0)
        check(name + ".freq", o.freq_indices, sel.last_frequency_indices)
        if expect is not None:
            assert p.tolist() == expect, (name, p.tolist(), expect)
        out[name] = dict(x=x, k=k, pmax=pmax, mpt=mpt, periods=p, amps=a, freq=sel.last_frequency_indices)
    torch.save(out, OUT / "selector_small.pt")
    print("selector_small:", {k: v["periods"].tolist() for k, v in out.items()})

    # BASELINE shapes: inputs regenerate from seeds, only outputs are stored
    big = {}
    shapes = {"etth1": (256, 96, 64), "elec": (64, 336, 128), "traffic": (32, 720, 256), "recursive": (512, 28, 128)}
    for wname, (B, L, C) in shapes.items():
        for kind in ("planted", "white"):
            x32 = syn.planted_features(B, L, C, seed=0) if kind == "planted" else syn.white_features(B, L, C, seed=1)
            for dt in (torch.float32, torch.bfloat16):
                x = x32.to(dt)
                k = 2 if wname == "recursive" else 5
                mpt = 7 if wname == "recursive" else 1
                sel = ref.FFTPeriodSelector(k_periods=k, pmax=L, min_period_threshold=mpt)
                p, a = sel(x)
                o = orc.select_periods(x, k, L, mpt)
                check(f"{wname}.{kind}.{dt}.periods", o.periods, p)
                check(f"{wname}.{kind}.{dt}.amps", o.amplitudes, a, 0.0)
                mean64 = orc.spectrum_float64(x)
                big[f"{wname}.{kind}.{'bf16' if dt == torch.bfloat16 else 'f32'}"] = dict(
                    B=B, L=L, C=C, k=k, mpt=mpt, periods=p, freq=sel.last_frequency_indices,
                    amps=a.float(), amp_mean=o.amp_mean, gap=orc.score_gap(mean64, k))
    torch.save(big, OUT / "selector_baseline_shapes.pt")
    print("selector_baseline_shapes:", {k: (v["periods"].tolist(), f"gap={v['gap']:.2e}") for k, v in big.items()})


def grouper_cases():
    out = {}
    g = torch.Generator().manual_seed(11)
    specs = {
        "dups_4448": ([4, 4, 4, 8], torch.tensor([[1.2, -0.7, 0.3, 0.1]]).expand(2, -1).contiguous(), 24, None, None, {}),
        "dups_4484": ([4, 4, 8, 4], torch.tensor([[0.5, -1.2, 0.3, 0.7]]).expand(2, -1).contiguous(), 25, None, None, {}),
        "too_long": ([64, 4], torch.tensor([[3.0, -1.0]]), 16, None, None, {}),
        "mixed_pad_L37": ([4, 5, 6, 11, 13], torch.randn(3, 5, generator=g), 37, 1, 37, {}),
        "invalid": ([0, -1], torch.tensor([[1.0, 1.0]]), 5, None, None, {}),
        "bounds": ([3, 9, 30, 12], torch.randn(2, 4, generator=g), 48, 4, 24, {}),
        "binning_maxuniq": ([3, 4, 6, 12], torch.tensor([[0.5, -0.2, 1.0, -1.5], [1.3, 0.1, -0.4, -2.0]]), 48, 1, 48,
                            {"TIMES_PERIOD_MAX_UNIQ": "2", "TIMES_PERIOD_BINNING": "log:2"}),
        "maxuniq_only": ([3, 4, 6, 12], torch.tensor([[0.5, -0.2, 1.0, -1.5]]), 48, None, None,
                         {"TIMES_PERIOD_MAX_UNIQ": "2"}),
    }
    for name, (per, amp, L, lo, hi, env) in specs.items():
        for k_, v_ in env.items():
            os.environ[k_] = v_
        r = ref.PeriodGrouper(torch.tensor(per, dtype=torch.long), amp, L, min_period=lo, max_period=hi).group()
        for k_ in env:
            os.environ.pop(k_)
        base = ref._resolve_log_binning_base(env.get("TIMES_PERIOD_BINNING"), None)
        mu = ref._resolve_scheduled_int(env.get("TIMES_PERIOD_MAX_UNIQ"), None)
        o = orc.group_periods(per, amp, L, lo, hi, base, mu)
        assert o.periods == r.periods.tolist() and o.pads == r.pad_lengths.tolist(), name
        assert o.cycles == r.cycles.tolist() and o.mapping == r.mapping.tolist(), name
        assert o.canonical == r.canonical_indices.tolist(), name
        check(name + ".logits", o.logits, r.logits)
        out[name] = dict(periods_in=per, amps=amp, L=L, lo=lo, hi=hi, env=env, periods=r.periods,
                         pads=r.pad_lengths, cycles=r.cycles, mapping=r.mapping, logits=r.logits,
                         canonical=r.canonical_indices)
    torch.save(out, OUT / "grouper.pt")
    print("grouper:", {k: (v["periods"].tolist(), v["mapping"].tolist()) for k, v in out.items()})


def _ref_block(wl: syn.Workload, w, layer=0, act="gelu"):
    blk = ref.TimesBlock(d_model=wl.d_model, kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0,
                         activation=act, d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio)
    pre = f"blocks.{layer}.inception."
    blk.inception.load_state_dict({k[len(pre):]: v for k, v in w.items() if k.startswith(pre)}, strict=True)
    return blk.eval()


def inception_and_block_cases():
    out = {}
    variants = {
        "toy": syn.WORKLOADS["toy"],
        "toy_ratio1": syn.Workload("toy_ratio1", 2, 24, 3, 6, 8, 1, 2, "f32", d_ff=8, kernel_set=((3, 3),),
                                   bottleneck_ratio=1.0),
        "toy_rect": syn.Workload("toy_rect", 2, 30, 3, 6, 12, 1, 3, "f32", d_ff=20, kernel_set=((1, 3), (5, 3)),
                                 bottleneck_ratio=2.0),
    }
    for vname, wl in variants.items():
        for act in ("gelu", "relu"):
            w = syn.stack_weights(wl, seed=3)
            blk = _ref_block(wl, w, 0, act)
            g = torch.Generator().manual_seed(5)
            grid = torch.randn(2, wl.d_model, 4, 6, generator=g)
            y_ref = blk.inception(grid)
            y_orc = orc.inception_stack(grid, w, "blocks.0.inception.", act)
            check(f"inception.{vname}.{act}", y_orc, y_ref)
            # block with a fixed selector (softmax mass spread over groups)
            x = syn.white_features(wl.B, wl.T, wl.d_model, seed=4)
            per = [5, 4, 4, 7, 60] if wl.T >= 30 else [3, 4, 4, 5, 40]
            amp = torch.randn(wl.B, len(per), generator=g)
            object.__setattr__(blk, "period_selector", FixedSelector(per, amp))
            b_ref = blk(x)
            tr = orc.timesblock_from_periods(x, per, amp, w, "blocks.0.inception.", act)
            check(f"block_fixed.{vname}.{act}", tr.out, b_ref)
            # block with the real FFT selector
            sel = ref.FFTPeriodSelector(wl.k_periods, wl.T, wl.min_period_threshold)
            object.__setattr__(blk, "period_selector", sel)
            f_ref = blk(x)
            tf = orc.timesblock_forward(x, w, "blocks.0.inception.", wl.k_periods, wl.T, wl.min_period_threshold, act)
            check(f"block_fft.{vname}.{act}", tf.out, f_ref)
            out[f"{vname}.{act}"] = dict(workload=wl.as_dict(), weight_seed=3, grid=grid, inception_out=y_ref, x=x,
                                         fixed_periods=per, fixed_amps=amp, block_fixed_out=b_ref,
                                         fixed_weights=tr.weights, fixed_group_periods=tr.groups.periods,
                                         block_fft_out=f_ref, fft_periods=sel.last_selected_periods,
                                         fft_deltas=[d.clone() for d in tf.deltas])
    torch.save(out, OUT / "block_toy.pt")
    print("block_toy:", {k: v["fft_periods"].tolist() for k, v in out.items()})


def _ref_stack(wl, w, x):
    """Reference TimesBlock stack + shared LayerNorm exactly as TimesNet.forward runs it."""
    sel = ref.FFTPeriodSelector(wl.k_periods, wl.T, wl.min_period_threshold)
    ln = torch.nn.LayerNorm(wl.d_model)
    ln.load_state_dict({"weight": w["layer_norm.weight"], "bias": w["layer_norm.bias"]})
    seq = x
    periods, outs = [], []
    for i in range(wl.n_layers):
        blk = _ref_block(wl, w, i)
        blk.block_index = i
        object.__setattr__(blk, "period_selector", sel)
        upd = blk(seq)
        periods.append(sel.last_selected_periods.clone())
        outs.append(upd)
        seq = ref._apply_norm_module(ln, seq + (upd - seq))
    return seq, periods, outs


def stack_cases():
    out = {}
    for wname, B in (("toy", None), ("toy_bf16", None), ("mid", None), ("etth1", 8), ("elec", 2)):
        wl = syn.WORKLOADS[wname]
        if B is not None:
            wl = syn.Workload(**{**wl.__dict__, "B": B})
        w = syn.stack_weights(wl, seed=0)
        for kind in ("planted", "white"):
            x = (syn.planted_features(wl.B, wl.T, wl.d_model, 0) if kind == "planted"
                 else syn.white_features(wl.B, wl.T, wl.d_model, 1))
            for dname in (("f32", "bf16") if B is not None else (wl.dtype,)):
                xd = x.to(syn.torch_dtype(dname))
                with torch.no_grad():
                    y_ref, periods, outs = _ref_stack(wl, w, xd)
                    trace = []
                    y_orc = orc.stack_forward(xd, w, wl.n_layers, wl.k_periods, wl.T, wl.min_period_threshold,
                                              trace=trace)
                d = check(f"stack.{wname}.{kind}.{dname}", y_orc, y_ref, 1e-5 if dname == "f32" else 1.6e-2)
                full = wl.B * wl.T * wl.d_model <= 8192
                out[f"{wname}.{kind}.{dname}"] = dict(
                    workload=wl.as_dict(), weight_seed=0, input=kind, periods=[p.tolist() for p in periods],
                    block0_weights=trace[0].weights.float(),
                    out_full=y_ref.float() if full else None, out_sub=subsample(y_ref),
                    block0_out_sub=subsample(outs[0]), out_abs_mean=y_ref.float().abs().mean().item(),
                    oracle_vs_ref=d)
    torch.save(out, OUT / "stack.pt")
    print("stack:", {k: v["periods"] for k, v in out.items()})


def _build_ref_model(wl: syn.Workload, with_static: bool, x, static, ids, min_sigma_vector=None):
    torch.manual_seed(0)
    m = ref.TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers,
                     k_periods=wl.k_periods, kernel_set=[list(k) for k in wl.kernel_set], dropout=0.1,
                     activation="gelu", mode=wl.mode, d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio,
                     min_period_threshold=wl.min_period_threshold, use_checkpoint=False,
                     use_zero_mean_context=wl.context_rank > 0, context_rank=wl.context_rank,
                     context_scale=0.05, static_proj_dim=6 if with_static else None,
                     min_sigma_vector=min_sigma_vector)
    with torch.no_grad():
        m(x[:1], series_static=static, series_ids=ids)       # lazy build (SURVEY 8c protocol step 1)
    m.eval()
    sd = syn.reseed_module_state(m, seed=9)
    if min_sigma_vector is not None:
        sd["min_sigma_vector"] = m.min_sigma_vector.clone()
    m.load_state_dict(sd, strict=True)
    return m, sd


def model_cases():
    out = {}
    specs = {
        "toy_direct": (syn.Workload("toy_direct", 3, 48, 5, 12, 16, 2, 3, "f32", d_ff=32, context_rank=4), True, 48),
        "toy_longhist": (syn.Workload("toy_longhist", 2, 48, 4, 60, 16, 1, 3, "f32", d_ff=32), False, 56),
        "toy_recursive": (syn.Workload("toy_recursive", 6, 28, 1, 4, 16, 2, 2, "f32", d_ff=32,
                                       min_period_threshold=7, mode="recursive", context_rank=4), True, 28),
    }
    for name, (wl, with_static, T_in) in specs.items():
        g = torch.Generator().manual_seed(21)
        x = syn.planted_series(wl.B, T_in, wl.N, seed=3)
        static = torch.randn(wl.N, 5, generator=g) if with_static else None
        ids = torch.arange(wl.N)
        msv = (0.01 + 0.1 * torch.rand(wl.N, generator=g)) if name == "toy_direct" else None
        m, sd = _build_ref_model(wl, with_static, x, static, ids, msv)
        cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu",
                           wl.min_period_threshold, 1e-3, wl.context_rank > 0, wl.context_rank)
        with torch.no_grad():
            r_ref, d_ref = m(x, series_static=static, series_ids=ids)
            r_orc, d_orc = orc.timesnet_forward(x, sd, cfg, series_static=static, series_ids=ids,
                                                min_sigma_vector=msv)
        check(name + ".rate", r_orc, r_ref)
        check(name + ".disp", d_orc, d_ref)
        y = syn.poisson_targets(wl.B, r_ref.shape[1], wl.N, 5.0, seed=2)
        mask = (torch.rand(y.shape, generator=g) > 0.2)
        nll_ref = ref_nll(y, r_ref, d_ref, mask)
        check(name + ".nll", orc.nb_nll(y, r_orc, d_orc, mask), nll_ref)
        rec = {}
        if wl.mode == "recursive":
            with torch.no_grad():
                rr, rd = ref_recursive(m, x, wl.H, series_static=static, series_ids=ids)
                orr, ord_ = orc.forecast_recursive(x, wl.H, sd, cfg, series_static=static, series_ids=ids)
            check(name + ".rec_rate", orr, rr, 2e-5)
            check(name + ".rec_disp", ord_, rd, 2e-5)
            rec = dict(rec_rate=rr, rec_disp=rd)
        out[name] = dict(workload=wl.as_dict(), T_in=T_in, x_seed=3, state_seed=9, with_static=with_static,
                         static=static, ids=ids, min_sigma_vector=msv, state={k: v.clone() for k, v in sd.items()},
                         rate=r_ref, disp=d_ref, y=y, mask=mask, nll=nll_ref, **rec)
    # context op on its own (LowRankTemporalContext, reference tests have none)
    g = torch.Generator().manual_seed(5)
    coeff = torch.randn(3, 7, 6, generator=g)
    tc = ref.LowRankTemporalContext(rank=6, init_scale=0.3)
    c_ref = tc(coeff, 40)
    check("lowrank", orc.lowrank_context(coeff, 40, torch.tensor(0.3)), c_ref)
    out["lowrank"] = dict(coeff=coeff, length=40, scale=0.3, ctx=c_ref.detach())
    # NLL known-answer (reference tests/test_negative_binomial_nll.py style)
    y = torch.tensor([[0.0, 1.0, 5.0], [float("nan"), 2.0, 30.0]])
    rate = torch.tensor([[0.5, 1.5, 4.0], [1.0, 1e-9, 25.0]])
    disp = torch.tensor([[0.3, 1e-9, 2.0], [1.0, 0.7, 0.05]])
    mask = torch.tensor([[1, 1, 0], [1, 1, 1]])
    n_ref = ref_nll(y, rate, disp, mask)
    check("nll_small", orc.nb_nll(y, rate, disp, mask), n_ref)
    # the reference multiplies ll by a 0/1 weight, so a NaN target poisons the sum even when masked
    assert bool(torch.isnan(n_ref))
    y2 = torch.nan_to_num(y, nan=3.0)
    n2 = ref_nll(y2, rate, disp, mask)
    check("nll_small_finite", orc.nb_nll(y2, rate, disp, mask), n2)
    out["nll_small"] = dict(y=y, y_finite=y2, rate=rate, disp=disp, mask=mask, nll_with_nan=n_ref, nll=n2,
                            nll_nomask=ref_nll(y2, rate, disp))
    torch.save(out, OUT / "model_toy.pt")
    print("model_toy:", {k: float(v["nll"]) for k, v in out.items() if "nll" in v})


def mean_vs_sum_check():
    """The oracle writes mean(dim=0) as sum/B (shardable); verify bitwise equality on the bench inputs."""
    for (B, L, C) in ((256, 96, 64), (64, 336, 128), (32, 720, 256)):
        med = orc.channel_median_spectrum(syn.planted_features(B, L, C, 0))
        assert torch.equal(med.mean(dim=0), med.sum(dim=0) / float(B)), (B, L, C)
    print("mean == sum/B on bench inputs: ok")


if __name__ == "__main__":
    with torch.no_grad():
        mean_vs_sum_check()
        selector_cases()
        grouper_cases()
        inception_and_block_cases()
        stack_cases()
        model_cases()
    tot = sum(p.stat().st_size for p in OUT.glob("*.pt"))
    print(f"golden fixtures written to {OUT} ({tot / 1e6:.2f} MB)")
