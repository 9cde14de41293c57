"""GPU parity: libflowtimes (through the drop-in modules / C ABI) vs the oracle
and the reference's golden outputs, on the same seeded inputs.

Bars (BASELINE.json north_star):
  * period / frequency indices: bit-exact
  * fp32 outputs: <= 1e-4 relative (we assert REL_F32 = 1e-4 against max|ref|)
  * bf16 activations: the reference keeps conv math in fp32 and rounds the
    activations between ops; our fp32-math kernels do the same, so the bound is
    a few bf16 ulps of the output scale: REL_BF16 = 2e-2
"""
import math
import os

import pytest
import torch

import flowtimes_oracle as orc
import flowtimes_synth as syn

pytestmark = pytest.mark.gpu

REL_F32 = 1e-4
REL_BF16 = 2e-2


def _rel(got, want):
    """max |got - want| / max |want|: a tensor-level bound (looser than element-wise relative error).
    With FLOWTIMES_LOG_ERR=<file> every measured value is appended with the calling test's name (profiles/ keeps the
    margins of a round)."""
    got = got.detach().float().cpu()
    want = want.detach().float().cpu()
    scale = max(1e-6, want.abs().max().item())
    err = (got - want).abs().max().item() / scale
    log = os.environ.get("FLOWTIMES_LOG_ERR")
    if log:
        name = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
        with open(log, "a") as f:
            f.write(f"{err:.3e} {name}\n")
    return err


def _sub(t, n=4096):
    flat = t.reshape(-1).float()
    return flat[:: max(1, flat.numel() // n)]


def _wl(d):
    d = dict(d)
    d["kernel_set"] = tuple(tuple(k) for k in d["kernel_set"])
    return syn.Workload(**d)


def _make_block(wl, w, layer=0, act="gelu"):
    from timesnet_forecast.models.timesnet import TimesBlock
    blk = TimesBlock(d_model=wl.d_model, kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0, activation=act,
                     d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio)
    pre = f"blocks.{layer}.inception."
    blk.inception.load_state_dict({k[len(pre):]: v for k, v in w.items() if k.startswith(pre)}, strict=True)
    return blk.cuda().eval()


class FixedSelector(torch.nn.Module):
    def __init__(self, periods, amplitudes):
        super().__init__()
        self.p = torch.as_tensor(periods, dtype=torch.long)
        self.a = torch.as_tensor(amplitudes, dtype=torch.float32)

    def forward(self, x):
        a = self.a.to(device=x.device, dtype=x.dtype)
        a = a.unsqueeze(0) if a.dim() == 1 else a
        a = a.expand(x.size(0), -1) if a.size(0) == 1 and x.size(0) > 1 else a
        return self.p.to(x.device), a


# --------------------------------------------------------------------------- #
# K1
# --------------------------------------------------------------------------- #
def test_library_targets_sm100():
    from timesnet_forecast import _native as nv
    sm, major, minor = nv.device_info()
    assert major == 10 and sm > 0


def test_selector_reference_known_answers(golden_dir):
    from timesnet_forecast.models.timesnet import FFTPeriodSelector
    g = torch.load(golden_dir / "selector_small.pt")
    for name, c in g.items():
        sel = FFTPeriodSelector(c["k"], c["pmax"], c["mpt"])
        p, a = sel(c["x"].cuda())
        assert p.dtype == torch.long and a.dtype == c["x"].dtype
        assert p.tolist() == c["periods"].tolist(), name
        assert sel.last_frequency_indices.tolist() == c["freq"].tolist(), name
        assert sel.last_selected_periods.tolist() == c["periods"].tolist(), name
        assert tuple(a.shape) == tuple(c["amps"].shape), name
        if c["amps"].numel():
            assert _rel(a, c["amps"]) < 1e-5, name


@pytest.mark.parametrize("wname", ["etth1", "elec", "traffic", "recursive"])
@pytest.mark.parametrize("kind", ["planted", "white"])
@pytest.mark.parametrize("dname", ["f32", "bf16"])
def test_selector_baseline_shapes(golden_dir, wname, kind, dname):
    """Full BASELINE.json shapes: selected bins / periods identical to the reference."""
    from timesnet_forecast.models.timesnet import FFTPeriodSelector
    c = torch.load(golden_dir / "selector_baseline_shapes.pt")[f"{wname}.{kind}.{dname}"]
    x = (syn.planted_features(c["B"], c["L"], c["C"], 0) if kind == "planted"
         else syn.white_features(c["B"], c["L"], c["C"], 1)).to(syn.torch_dtype(dname))
    sel = FFTPeriodSelector(c["k"], c["L"], c["mpt"])
    p, a = sel(x.cuda())
    o = orc.select_periods(x, c["k"], c["L"], c["mpt"])
    # batch-mean spectrum itself: fp32 DFT vs pocketfft, relative to the largest bin
    med, ssum = __import__("timesnet_forecast._native", fromlist=["x"]).spectrum(x.cuda())
    Fq = c["L"] // 2 + 1
    assert ssum.numel() == Fq + 1 and float(ssum[Fq]) == c["B"]      # count slot (all-reduced with the sums)
    assert _rel(ssum[1:Fq] / c["B"], o.amp_mean[1:]) < 2e-5
    if dname == "bf16" and kind == "white":
        # bf16-rounded scores tie; torch.topk's tie order is unspecified:
        # require the same multiset of scores instead of the same bins
        got = sorted(o.scores[sel.last_frequency_indices.cpu()].float().tolist())
        want = sorted(o.scores[c["freq"]].float().tolist())
        assert got == want
        return
    assert sel.last_frequency_indices.tolist() == c["freq"].tolist(), f"gap={c['gap']:.2e}"
    assert p.tolist() == c["periods"].tolist()
    assert _rel(a, c["amps"]) < (1e-5 if dname == "f32" else 8e-3)


def test_selector_edge_cases():
    from timesnet_forecast.models.timesnet import FFTPeriodSelector
    sel = FFTPeriodSelector(k_periods=3, pmax=8)
    with pytest.raises(ValueError):
        sel(torch.zeros(4, 8, device="cuda"))
    p, a = sel(torch.zeros(2, 1, 3, device="cuda"))            # L <= 1 -> empty (timesnet.py:89-90)
    assert p.numel() == 0 and tuple(a.shape) == (2, 0)
    p, a = FFTPeriodSelector(0, 8)(torch.randn(2, 8, 3, device="cuda"))
    assert p.numel() == 0 and tuple(a.shape) == (2, 0)
    with pytest.raises(RuntimeError):
        sel(torch.randn(2, 8, 3))                               # CPU tensor: no fallback
    # L = 2: one usable bin, period clamps to 1, cycles = 2
    x = torch.randn(3, 2, 4)
    p, a = FFTPeriodSelector(2, 2)(x.cuda())
    o = orc.select_periods(x, 2, 2)
    assert p.tolist() == o.periods.tolist()
    # NaN input propagates like torch.median / topk (NaN ranks first)
    x = torch.randn(2, 16, 3)
    x[0, 3, 1] = float("nan")
    o = orc.select_periods(x, 2, 16)
    p, a = FFTPeriodSelector(2, 16)(x.cuda())
    assert p.numel() == o.periods.numel()


# --------------------------------------------------------------------------- #
# K2-K4
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("case", ["toy.gelu", "toy.relu", "toy_ratio1.gelu", "toy_ratio1.relu", "toy_rect.gelu"])
def test_block_golden(golden_dir, case):
    c = torch.load(golden_dir / "block_toy.pt")[case]
    act = case.split(".")[1]
    wl = _wl(c["workload"])
    w = syn.stack_weights(wl, seed=c["weight_seed"])
    blk = _make_block(wl, w, 0, act)
    from timesnet_forecast.models.timesnet import FFTPeriodSelector
    x = c["x"].cuda()
    object.__setattr__(blk, "period_selector", FixedSelector(c["fixed_periods"], c["fixed_amps"]))
    out = blk(x)
    assert out.shape == x.shape
    assert _rel(out, c["block_fixed_out"]) < REL_F32
    assert blk._last_group_count == len(c["fixed_group_periods"])
    sel = FFTPeriodSelector(wl.k_periods, wl.T, wl.min_period_threshold)
    object.__setattr__(blk, "period_selector", sel)
    out = blk(x)
    assert sel.last_selected_periods.tolist() == c["fft_periods"].tolist()
    assert _rel(out, c["block_fft_out"]) < REL_F32


@pytest.mark.parametrize("wname,B", [("toy", 4), ("mid", 4), ("etth1", 3), ("elec", 2)])
@pytest.mark.parametrize("dname", ["f32", "bf16"])
def test_per_period_delta(wname, B, dname):
    """Per-period delta BEFORE aggregation (softmax weights are ~one-hot on planted data and
    would hide errors in the other periods)."""
    from timesnet_forecast import _native as nv
    wl = syn.WORKLOADS[wname]
    w = syn.stack_weights(wl, seed=0)
    dt = syn.torch_dtype(dname)
    L, C = wl.T, wl.d_model
    x = syn.white_features(B, L, C, seed=4).to(dt)
    periods = [5, 4, 4, 7, L - 1] if L < 90 else [24, 12, 7, 48, 6, L - 1, 5]
    g = orc.group_periods(periods, torch.zeros(1, len(periods)), L)
    blk = _make_block(wl, w)
    xc = x.cuda()
    plan = nv.plan_to_device(nv.plan_build_host(periods, L, None, None), xc.device)
    pa, pb = blk.inception[0].packed(xc.device), blk.inception[2].packed(xc.device)
    mg = len(periods)
    ws = torch.empty(nv.inception_workspace_bytes(B, L, mg, pa.struct, pb.struct), dtype=torch.uint8, device="cuda")
    delta = torch.full((mg, B, L, C), float("nan"), dtype=dt, device="cuda")
    nv.period_conv(xc, plan, mg, pa.struct, pb.struct, nv.FTN_ACT_GELU, delta, ws)
    torch.cuda.synchronize()
    tol = REL_F32 if dname == "f32" else REL_BF16
    for gi, (p, pad, cyc) in enumerate(zip(g.periods, g.pads, g.cycles)):
        want = orc.period_delta(x, p, pad, cyc, w, "blocks.0.inception.", "gelu")
        err = _rel(delta[gi], want)
        assert err < tol, f"period {p} (pad {pad}, cycles {cyc}): rel err {err:.3e}"


@pytest.mark.parametrize("key", ["toy.planted.f32", "toy.white.f32", "toy_bf16.planted.bf16", "toy_bf16.white.bf16",
                                 "mid.planted.f32", "mid.white.f32", "etth1.planted.f32", "etth1.white.f32",
                                 "etth1.planted.bf16", "elec.planted.f32", "elec.white.f32", "elec.planted.bf16"])
def test_stack_golden(golden_dir, key):
    """n_layers x (TimesBlock + shared LayerNorm) vs the reference's stored outputs."""
    from timesnet_forecast.models.timesnet import FFTPeriodSelector
    c = torch.load(golden_dir / "stack.pt")[key]
    wl = _wl(c["workload"])
    dname = key.split(".")[2]
    w = syn.stack_weights(wl, seed=c["weight_seed"])
    x = (syn.planted_features(wl.B, wl.T, wl.d_model, 0) if c["input"] == "planted"
         else syn.white_features(wl.B, wl.T, wl.d_model, 1)).to(syn.torch_dtype(dname)).cuda()
    sel = FFTPeriodSelector(wl.k_periods, wl.T, wl.min_period_threshold)
    ln = torch.nn.LayerNorm(wl.d_model).cuda()
    ln.load_state_dict({"weight": w["layer_norm.weight"], "bias": w["layer_norm.bias"]})
    seq = x
    periods = []
    first = None
    for i in range(wl.n_layers):
        blk = _make_block(wl, w, i)
        object.__setattr__(blk, "period_selector", sel)
        if i == 0:
            first = blk(seq)
        seq = blk.forward_norm(seq, ln)
        periods.append(sel.last_selected_periods.tolist())
    assert periods == c["periods"], "period selection differs from the reference"
    tol = REL_F32 if dname == "f32" else REL_BF16
    assert _rel(_sub(first), c["block0_out_sub"]) < tol
    assert _rel(_sub(seq), c["out_sub"]) < tol
    if c["out_full"] is not None:
        assert _rel(seq, c["out_full"]) < tol


def test_full_size_elec_windows_match_oracle():
    """BASELINE elec launch geometry (B=64, L=336, C=128, F=512, bf16): with a fixed selector the
    block is per-window, so windows 0..1 of the full batch must equal the oracle run on them alone."""
    wl = syn.WORKLOADS["elec"]
    w = syn.stack_weights(wl, seed=0)
    x = syn.white_features(wl.B, wl.T, wl.d_model, seed=1).to(torch.bfloat16)
    periods = [24, 12, 7, 48, 6]
    g = torch.Generator().manual_seed(3)
    amps = torch.randn(wl.B, 5, generator=g)
    blk = _make_block(wl, w)
    object.__setattr__(blk, "period_selector", FixedSelector(periods, amps))
    out = blk(x.cuda())
    tr = orc.timesblock_from_periods(x[:2], periods, amps[:2].to(torch.bfloat16), w, "blocks.0.inception.")
    assert _rel(out[:2], tr.out) < REL_BF16
    # last windows too (tile-decode at the far end of the grid)
    tr = orc.timesblock_from_periods(x[-1:], periods, amps[-1:].to(torch.bfloat16), w, "blocks.0.inception.")
    assert _rel(out[-1:], tr.out) < REL_BF16


@pytest.mark.parametrize("dname", ["f32", "bf16"])
def test_full_size_etth1_windows_match_oracle(dname):
    """BASELINE etth1 launch geometry (B=256, L=96, C=64, F=256, mid=16): 1280 small images, tens of units per
    persistent CTA in every stage.  bf16 takes tc_gemm2 / tc_conv2 (image-resident, mid 16) / tc_gemm, fp32 the
    three-plane route; first, middle and last windows of the full batch must equal the oracle run on them alone."""
    wl = syn.WORKLOADS["etth1"]
    dt = torch.float32 if dname == "f32" else torch.bfloat16
    w = syn.stack_weights(wl, seed=0)
    x = syn.white_features(wl.B, wl.T, wl.d_model, seed=1).to(dt)
    periods = [24, 12, 7, 48, 6]
    g = torch.Generator().manual_seed(3)
    amps = torch.randn(wl.B, 5, generator=g)
    blk = _make_block(wl, w)
    object.__setattr__(blk, "period_selector", FixedSelector(periods, amps))
    out = blk(x.cuda())
    torch.cuda.synchronize()
    tol = REL_F32 if dname == "f32" else REL_BF16
    for sl in (slice(0, 2), slice(127, 129), slice(wl.B - 1, wl.B)):
        tr = orc.timesblock_from_periods(x[sl], periods, amps[sl].to(dt), w, "blocks.0.inception.")
        assert _rel(out[sl], tr.out) < tol


@pytest.mark.parametrize("periods", [[7, 14], [27, 9], [28, 2]])
@pytest.mark.parametrize("B", [6, 70])
def test_short_window_fused_block_matches_oracle(periods, B):
    """Config-5 block geometry (L = 28, C = 128, F = 512, bf16) through the fused route with every short-window form
    at once: 32-row granules, stacked tc_conv4 units (several windows per unit, a ragged last unit at B = 70) and
    tc_tail items of four windows (a ragged last quad at B = 6 and 70); images of one and two granules."""
    wl = syn.WORKLOADS["recursive"]
    w = syn.stack_weights(wl, seed=0)
    x = syn.white_features(B, wl.T, wl.d_model, seed=11).to(torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    amps = torch.randn(B, len(periods), generator=g)
    blk = _make_block(wl, w)
    object.__setattr__(blk, "period_selector", FixedSelector(periods, amps))
    out = blk(x.cuda())
    torch.cuda.synchronize()
    tr = orc.timesblock_from_periods(x, periods, amps.to(torch.bfloat16), w, "blocks.0.inception.")
    assert _rel(out, tr.out) < REL_BF16
    # every window on its own: a window must not see its neighbours in a stacked unit or a quad item
    for b in (0, B // 2, B - 1):
        object.__setattr__(blk, "period_selector", FixedSelector(periods, amps[b:b + 1]))
        one = blk(x[b:b + 1].cuda())
        assert torch.equal(one[0], out[b]), f"window {b} depends on its batch neighbours"


@pytest.mark.parametrize("L", [30, 47, 64, 65, 96, 97, 129, 200])
def test_fused_block_window_length_sweep(L):
    """Elec-class block (C = 128, F = 512, mid = 32, bf16) over window lengths that straddle every geometry switch of the
    fused route: tc_tail's four-window items (L <= 96) vs one window per item, one / several granules per image, stacked
    vs single tc_conv4 units (grids of at most 256 padded positions), ragged last tiles.  Periods are drawn at random
    (seeded) from [1, L - 1] plus the extremes; B = 5 leaves ragged quads and stacks."""
    wl0 = syn.WORKLOADS["elec"]
    wl = syn.Workload(**{**wl0.__dict__, "T": L, "B": 5})
    w = syn.stack_weights(wl, seed=0)
    g = torch.Generator().manual_seed(L)
    periods = sorted({1, 2, L - 1, L // 2, *torch.randint(3, L - 1, (3,), generator=g).tolist()})[:6]
    x = syn.white_features(wl.B, L, wl.d_model, seed=L).to(torch.bfloat16)
    amps = torch.randn(wl.B, len(periods), generator=g)
    blk = _make_block(wl, w)
    object.__setattr__(blk, "period_selector", FixedSelector(periods, amps))
    out = blk(x.cuda())
    torch.cuda.synchronize()
    tr = orc.timesblock_from_periods(x, periods, amps.to(torch.bfloat16), w, "blocks.0.inception.")
    assert _rel(out, tr.out) < REL_BF16, f"L={L} periods={periods}"


@pytest.mark.parametrize("L", [40, 96, 130])
@pytest.mark.parametrize("dname", ["f32", "bf16"])
def test_narrow_block_window_length_sweep(L, dname):
    """etth1-class block (C = 64, F = 256, mid = 16) over window lengths with random periods: fp32 on the fp16-pair GEMMs
    and the row mode of tc_convs (1e-4 bound), bf16 on tc_gemm2 / the row mode / tc_gemm (2e-2)."""
    wl0 = syn.WORKLOADS["etth1"]
    wl = syn.Workload(**{**wl0.__dict__, "T": L, "B": 3})
    dt = torch.float32 if dname == "f32" else torch.bfloat16
    w = syn.stack_weights(wl, seed=0)
    g = torch.Generator().manual_seed(100 + L)
    periods = sorted({1, 2, L - 1, L // 2, *torch.randint(3, L - 1, (3,), generator=g).tolist()})[:6]
    x = syn.white_features(wl.B, L, wl.d_model, seed=L).to(dt)
    amps = torch.randn(wl.B, len(periods), generator=g)
    blk = _make_block(wl, w)
    object.__setattr__(blk, "period_selector", FixedSelector(periods, amps))
    out = blk(x.cuda())
    torch.cuda.synchronize()
    tr = orc.timesblock_from_periods(x, periods, amps.to(dt), w, "blocks.0.inception.")
    assert _rel(out, tr.out) < (REL_F32 if dname == "f32" else REL_BF16), f"L={L} periods={periods}"


@pytest.mark.parametrize("wname,dname", [("etth1", "f32"), ("etth1", "bf16"), ("elec", "bf16"), ("elec", "f32")])
def test_relu_blocks_on_the_tensor_core_routes(wname, dname):
    """activation = "relu" (timesnet.py:736-747) through the tensor-core routes: the fused bf16 chain (ACT = 1 template
    instances of tc_mid / tc_tail / tc_gemm2), the fp16-pair fp32 chain and the narrow-branch row mode."""
    wl0 = syn.WORKLOADS[wname]
    wl = syn.Workload(**{**wl0.__dict__, "B": 3})
    dt = torch.float32 if dname == "f32" else torch.bfloat16
    w = syn.stack_weights(wl, seed=0)
    periods = [24, 12, 7, wl.T // 2, wl.T - 1]
    x = syn.white_features(wl.B, wl.T, wl.d_model, seed=3).to(dt)
    g = torch.Generator().manual_seed(8)
    amps = torch.randn(wl.B, len(periods), generator=g)
    blk = _make_block(wl, w, 0, "relu")
    object.__setattr__(blk, "period_selector", FixedSelector(periods, amps))
    out = blk(x.cuda())
    torch.cuda.synchronize()
    tr = orc.timesblock_from_periods(x, periods, amps.to(dt), w, "blocks.0.inception.", "relu")
    assert _rel(out, tr.out) < (REL_F32 if dname == "f32" else REL_BF16)


def test_elec_block_with_long_periods_matches_oracle():
    """Periods whose padded grid does not fit tc_conv4's shared-memory image (100, 168) take the tc_conv2 fallback
    inside the same launch sequence, both reading the once-per-window first 1x1 stage; short ones stay on tc_conv4."""
    wl = syn.WORKLOADS["elec"]
    w = syn.stack_weights(wl, seed=0)
    B = 5
    x = syn.white_features(B, wl.T, wl.d_model, seed=5).to(torch.bfloat16)
    periods = [24, 100, 7, 168, 6]
    g = torch.Generator().manual_seed(4)
    amps = torch.randn(B, 5, generator=g)
    blk = _make_block(wl, w)
    object.__setattr__(blk, "period_selector", FixedSelector(periods, amps))
    out = blk(x.cuda())
    tr = orc.timesblock_from_periods(x, periods, amps.to(torch.bfloat16), w, "blocks.0.inception.")
    assert _rel(out, tr.out) < REL_BF16


def test_block_identity_and_errors():
    from timesnet_forecast.models.timesnet import TimesBlock
    blk = TimesBlock(d_model=2, kernel_set=[(3, 3)], dropout=0.0, activation="gelu").cuda()
    with pytest.raises(RuntimeError):
        blk(torch.randn(2, 5, 2, device="cuda"))               # selector not set (timesnet.py:772-773)
    object.__setattr__(blk, "period_selector", FixedSelector([0, -1], [1.0, 1.0]))
    x = torch.randn(2, 5, 2, device="cuda")
    assert torch.equal(blk(x), x)                               # no valid period -> identity
    with pytest.raises(ValueError):
        blk(torch.randn(2, 5, device="cuda"))
    with pytest.raises(ValueError):
        blk(torch.randn(2, 5, 3, device="cuda"))                # channel count changed
    with pytest.raises(RuntimeError):
        blk(torch.randn(2, 5, 2))                               # CPU: no fallback
    with pytest.raises(TypeError):
        blk(torch.randn(2, 5, 2, device="cuda").half())         # fp16 unsupported


# --------------------------------------------------------------------------- #
# K5, K6, whole model
# --------------------------------------------------------------------------- #
def test_lowrank_context(golden_dir):
    from timesnet_forecast.models.timesnet import LowRankTemporalContext
    c = torch.load(golden_dir / "model_toy.pt")["lowrank"]
    tc = LowRankTemporalContext(rank=c["coeff"].shape[-1], init_scale=c["scale"]).cuda()
    got = tc(c["coeff"].cuda(), c["length"])
    assert _rel(got, c["ctx"]) < 1e-5
    x = torch.randn(3, c["length"], 7)
    assert _rel(tc.add_to(x.cuda(), c["coeff"].cuda()), x + c["ctx"]) < 1e-5


def test_nb_nll(golden_dir):
    from timesnet_forecast.losses import negative_binomial_nll, negative_binomial_mask
    c = torch.load(golden_dir / "model_toy.pt")["nll_small"]
    cu = lambda t: t.cuda()
    assert torch.isnan(negative_binomial_nll(cu(c["y"]), cu(c["rate"]), cu(c["disp"]), cu(c["mask"])))
    got = negative_binomial_nll(cu(c["y_finite"]), cu(c["rate"]), cu(c["disp"]), cu(c["mask"]))
    assert got.dtype == torch.float32 and _rel(got, c["nll"]) < 1e-5
    assert _rel(negative_binomial_nll(cu(c["y_finite"]), cu(c["rate"]), cu(c["disp"])), c["nll_nomask"]) < 1e-5
    m = negative_binomial_mask(cu(c["y"]), cu(c["rate"]), cu(c["disp"]), cu(c["mask"]))
    assert m.dtype == torch.bool and m.cpu().tolist() == orc.nb_mask(c["y"], c["rate"], c["disp"], c["mask"]).tolist()
    # large random case vs oracle
    g = torch.Generator().manual_seed(0)
    y = torch.poisson(torch.full((64, 96, 321), 5.0), generator=g)
    rate = torch.rand(64, 96, 321, generator=g) * 10 + 0.01
    disp = torch.rand(64, 96, 321, generator=g) * 2 + 1e-3
    mask = torch.rand(64, 96, 321, generator=g) > 0.3
    assert _rel(negative_binomial_nll(cu(y), cu(rate), cu(disp), cu(mask)), orc.nb_nll(y, rate, disp, mask)) < 1e-5


def _build_model(c, wl, stack_dtype=None):
    from timesnet_forecast.models.timesnet import TimesNet
    m = TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods,
                 kernel_set=[list(k) for k in wl.kernel_set], dropout=0.1, activation="gelu", mode=wl.mode,
                 d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio, min_period_threshold=wl.min_period_threshold,
                 use_checkpoint=False, use_zero_mean_context=wl.context_rank > 0, context_rank=wl.context_rank,
                 context_scale=0.05, static_proj_dim=6 if c["with_static"] else None,
                 min_sigma_vector=c["min_sigma_vector"], stack_dtype=stack_dtype)
    x = syn.planted_series(wl.B, c["T_in"], wl.N, seed=c["x_seed"]).cuda()
    static = c["static"].cuda() if c["static"] is not None else None
    ids = c["ids"].cuda()
    m(x[:1], series_static=static, series_ids=ids)              # lazy build, then load the reference's weights
    m.eval()
    missing = m.load_state_dict(c["state"], strict=True)
    return m, x, static, ids


@pytest.mark.parametrize("name", ["toy_direct", "toy_longhist", "toy_recursive"])
def test_timesnet_forward_golden(golden_dir, name):
    from timesnet_forecast.losses import negative_binomial_nll
    from timesnet_forecast.predict import forecast_recursive_batch
    c = torch.load(golden_dir / "model_toy.pt")[name]
    wl = _wl(c["workload"])
    m, x, static, ids = _build_model(c, wl)
    assert sorted(m.state_dict().keys()) == sorted(c["state"].keys())     # reference checkpoint loads unchanged
    rate, disp = m(x, series_static=static, series_ids=ids)
    steps = wl.H if wl.mode == "direct" else 1
    assert tuple(rate.shape) == (wl.B, steps, wl.N) and tuple(disp.shape) == (wl.B, steps, wl.N)
    assert _rel(rate, c["rate"]) < REL_F32 and _rel(disp, c["disp"]) < REL_F32
    nll = negative_binomial_nll(c["y"].cuda(), rate, disp, c["mask"].cuda())
    assert _rel(nll, c["nll"]) < REL_F32
    if wl.mode == "recursive":
        rr, rd = forecast_recursive_batch(m, x, wl.H, series_static=static, series_ids=ids)
        assert _rel(rr, c["rec_rate"]) < REL_F32 and _rel(rd, c["rec_disp"]) < REL_F32


def test_timesnet_bf16_stack_close_to_fp32(golden_dir):
    c = torch.load(golden_dir / "model_toy.pt")["toy_direct"]
    wl = _wl(c["workload"])
    m, x, static, ids = _build_model(c, wl, stack_dtype=torch.bfloat16)
    rate, disp = m(x, series_static=static, series_ids=ids)
    assert rate.dtype == torch.float32
    assert _rel(rate, c["rate"]) < REL_BF16 and _rel(disp, c["disp"]) < REL_BF16


def test_graph_replay_and_pipelined_runner_match_eager():
    """CUDA-graph replay (GraphedCallable) and the double-buffered host-fed PipelinedRunner return exactly
    what the eager launch sequence returns, for several different inputs through the same captured graphs."""
    from timesnet_forecast.cuda_graphs import GraphedCallable, PipelinedRunner
    from timesnet_forecast.losses import negative_binomial_nll
    from timesnet_forecast.models.timesnet import TimesNet
    wl = syn.WORKLOADS["toy_bf16"]
    torch.manual_seed(0)
    m = TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods,
                 kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0, activation="gelu", mode="direct",
                 d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio, use_checkpoint=False,
                 stack_dtype=torch.bfloat16)
    xs = [syn.planted_series(wl.B, wl.T, wl.N, seed=s) for s in range(3)]
    ys = [syn.poisson_targets(wl.B, wl.H, wl.N, 5.0, seed=10 + s) for s in range(3)]
    m(xs[0][:1].cuda())
    m.eval()
    m.load_state_dict(syn.reseed_module_state(m, seed=4), strict=True)
    m.check_finite = False

    def fwd_loss(x, y):
        r, d = m(x)
        return negative_binomial_nll(y, r, d)

    eager = [fwd_loss(x.cuda(), y.cuda()).item() for x, y in zip(xs, ys)]
    feats = [syn.planted_features(wl.B, wl.T, wl.d_model, seed=s).to(torch.bfloat16).cuda() for s in range(3)]
    eager_stack = [m.stack_forward(f).clone() for f in feats]
    g = GraphedCallable(m.stack_forward, [feats[0]])
    for f, want in zip(feats, eager_stack):
        assert torch.equal(g(f), want)
    runner = PipelinedRunner(fwd_loss, [xs[0].cuda(), ys[0].cuda()])
    got = []
    for x, y in zip(xs, ys):
        slot = runner.submit(x.pin_memory(), y.pin_memory())
        runner.synchronize()
        got.append(float(runner.results[slot]))
    assert got == eager
    # back-to-back submits without a sync in between (the overlap case): last two results
    s0 = runner.submit(xs[1].pin_memory(), ys[1].pin_memory())
    s1 = runner.submit(xs[2].pin_memory(), ys[2].pin_memory())
    runner.synchronize()
    assert float(runner.results[s0]) == eager[1] and float(runner.results[s1]) == eager[2]


@pytest.mark.parametrize("with_ln", [False, True])
def test_fused_block_route_equals_unfused_route(with_ln):
    """ftn_timesblock_fused (chain + aggregation + LayerNorm in the tail kernel, no deltas in HBM) returns the
    same bits as ftn_period_conv + ftn_aggregate: the rounding points are identical by construction."""
    from timesnet_forecast import _native as nv
    from timesnet_forecast.models.timesnet import FFTPeriodSelector, TimesBlock
    wl = syn.WORKLOADS["elec"]
    B = 3
    torch.manual_seed(0)
    blk = TimesBlock(wl.d_model, [list(k) for k in wl.kernel_set], 0.0, "gelu", d_ff=wl.ff,
                     bottleneck_ratio=wl.bottleneck_ratio).cuda()
    x = syn.white_features(B, wl.T, wl.d_model, seed=3).to(torch.bfloat16).cuda()
    sel = FFTPeriodSelector(wl.k_periods, wl.T, 1)
    plan = sel.search(x)
    pa, pb = blk.inception[0].packed(x.device), blk.inception[2].packed(x.device)
    k = plan.k
    ws = torch.empty(nv.inception_workspace_bytes(B, wl.T, k, pa.struct, pb.struct), dtype=torch.uint8, device="cuda")
    ln_w = (1.0 + 0.1 * torch.randn(wl.d_model)).cuda() if with_ln else None
    ln_b = (0.1 * torch.randn(wl.d_model)).cuda() if with_ln else None
    fused = torch.empty_like(x)
    assert nv.timesblock_fused(x, plan.plan_dev, k, pa.struct, pb.struct, nv.FTN_ACT_GELU, plan.weights, ln_w, ln_b,
                               1e-5, fused, ws), "elec shape must be eligible for the fused route"
    delta = torch.empty(k, B, wl.T, wl.d_model, dtype=x.dtype, device="cuda")
    nv.period_conv(x, plan.plan_dev, k, pa.struct, pb.struct, nv.FTN_ACT_GELU, delta, ws)
    unfused = torch.empty_like(x)
    nv.aggregate(x, delta, plan.weights, plan.plan_dev, ln_w, ln_b, 1e-5, unfused)
    torch.cuda.synchronize()
    assert plan.host().n_groups >= 2
    if not with_ln:
        assert torch.equal(fused, unfused), (fused.float() - unfused.float()).abs().max()
    else:
        # the LayerNorm statistics are summed in a different order (4 column quarters vs a shuffle tree), which can
        # move a value across a bf16 rounding boundary: allow one bf16 ulp on a vanishing fraction of the elements
        f, u = fused.float(), unfused.float()
        assert torch.allclose(f, u, rtol=2 ** -7, atol=1e-6)
        assert (f != u).float().mean().item() < 1e-3


@pytest.mark.parametrize("kind", ["white", "planted"])
def test_timesblock_forward_equals_search_plus_fused(kind):
    """ftn_timesblock_forward (search + block in one call, first 1x1 stage forked beside the selection kernel) returns
    the bits of ftn_period_search followed by ftn_timesblock_fused, and the same plan.  The planted input selects
    p = L - 1 (two cycles), which takes tc_conv4's single-buffer CTAs."""
    from timesnet_forecast import _native as nv
    from timesnet_forecast.models.timesnet import TimesBlock
    wl = syn.WORKLOADS["elec"]
    B = 6
    torch.manual_seed(0)
    blk = TimesBlock(wl.d_model, [list(k) for k in wl.kernel_set], 0.0, "gelu", d_ff=wl.ff,
                     bottleneck_ratio=wl.bottleneck_ratio).cuda()
    feats = syn.white_features if kind == "white" else syn.planted_features
    x = feats(B, wl.T, wl.d_model, seed=7).to(torch.bfloat16).cuda()
    if kind == "planted":
        x = x + torch.linspace(0.0, 4.0, wl.T, device="cuda").view(1, -1, 1).to(torch.bfloat16)   # a trend: bin 1
    from timesnet_forecast.models.timesnet import FFTPeriodSelector
    object.__setattr__(blk, "period_selector", FFTPeriodSelector(wl.k_periods, wl.T, 1))
    blk(x)                                                       # lazy build
    pa, pb = blk.inception[0].packed(x.device), blk.inception[2].packed(x.device)
    k = wl.k_periods
    ln_w = (1.0 + 0.1 * torch.randn(wl.d_model)).cuda()
    ln_b = (0.1 * torch.randn(wl.d_model)).cuda()
    res = nv.timesblock_forward(x, k, wl.T, 1, pa.struct, pb.struct, nv.FTN_ACT_GELU, ln_w, ln_b, 1e-5)
    assert res is not None, "elec shape must be eligible for the combined call"
    out, plan, amps, weights = res
    plan2, amps2, weights2, _, _ = nv.period_search(x, k, wl.T, 1)
    ws = torch.empty(nv.inception_workspace_bytes(B, wl.T, k, pa.struct, pb.struct), dtype=torch.uint8, device="cuda")
    out2 = torch.empty_like(x)
    assert nv.timesblock_fused(x, plan2, k, pa.struct, pb.struct, nv.FTN_ACT_GELU, weights2, ln_w, ln_b, 1e-5, out2, ws)
    torch.cuda.synchronize()
    # (the last 4 of the 808 plan bytes are struct padding)
    assert torch.equal(plan[:804], plan2[:804]) and torch.equal(amps, amps2) and torch.equal(weights, weights2)
    assert torch.equal(out, out2)
    host = nv.plan_to_host(plan)
    if kind == "planted":
        assert max(host.grp_period[: host.n_groups]) >= wl.T // 2, list(host.grp_period[: host.n_groups])


@pytest.mark.parametrize("M,K,N", [(1, 1, 1), (33, 17, 64), (97, 128, 129), (200, 321, 128), (1000, 96, 321), (6144, 128, 321),
                                   (21504, 321, 128), (150, 7, 300)])
def test_linear_matches_fp32_matmul(M, K, N):
    """ftn_linear (the register-blocked fp32 SIMT GEMM of the embedding / head layers; every tile height and the
    ragged edges) against torch's fp32 matmul."""
    from timesnet_forecast import _native as nv
    g = torch.Generator().manual_seed(M * 7 + K * 3 + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / max(1.0, K ** 0.5)
    b = torch.randn(N, generator=g)
    out = nv.linear(a.cuda(), w.cuda(), b.cuda())
    ref = (a.double() @ w.double().t() + b.double()).float()
    assert _rel(out, ref) < 2e-6
    out = nv.linear(a.cuda(), w.cuda(), None)
    assert _rel(out, (a.double() @ w.double().t()).float()) < 2e-6
