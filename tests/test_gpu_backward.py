"""GPU: backward pass, first slice -- the native backward kernels behind timesnet_forecast/autograd.py
against float64 autograd of the reference formulas (losses.py:27-58, timesnet.py:2063-2093, F.layer_norm)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(got, want):
    want = want.detach().double().cpu()
    return (got.detach().double().cpu() - want).abs().max().item() / max(1e-12, want.abs().max().item())


def _nll64(y, rate, disp, mask, eps=1e-8):
    y = torch.clamp(y, min=0.0)
    a = torch.clamp(disp, min=eps)
    mu = torch.clamp(rate, min=eps)
    l1p = torch.log1p(a * mu)
    inv = 1.0 / a
    ll = (torch.lgamma(y + inv) - torch.lgamma(inv) - torch.lgamma(y + 1.0) - inv * l1p
          + y * (torch.log(a) + torch.log(mu) - l1p))
    w = (torch.isfinite(y) & torch.isfinite(mu) & torch.isfinite(a))
    if mask is not None:
        w = w & mask
    w = w.double()
    return -(ll * w).sum() / torch.clamp(w.sum(), min=1.0)


@pytest.mark.parametrize("with_mask", [False, True])
def test_nb_nll_backward_matches_float64_autograd(with_mask):
    from timesnet_forecast.losses import negative_binomial_nll
    g = torch.Generator().manual_seed(0)
    y = torch.poisson(torch.full((16, 12, 37), 5.0), generator=g)
    y[0, 0, :5] = torch.tensor([0.0, 1.0, 300.0, 2.5, 70.0])          # non-integer and large counts: digamma branch
    rate = torch.rand(16, 12, 37, generator=g) * 10 + 0.05
    disp = torch.rand(16, 12, 37, generator=g) * 2 + 1e-3
    disp[1, 0, :4] = torch.tensor([1e-3, 5e-3, 0.5, 30.0])
    mask = (torch.rand(16, 12, 37, generator=g) > 0.3) if with_mask else None
    r64 = rate.double().requires_grad_(True)
    d64 = disp.double().requires_grad_(True)
    loss64 = _nll64(y.double(), r64, d64, mask)
    loss64.backward()
    rc = rate.cuda().requires_grad_(True)
    dc = disp.cuda().requires_grad_(True)
    loss = negative_binomial_nll(y.cuda(), rc, dc, None if mask is None else mask.cuda())
    assert abs(loss.item() - loss64.item()) / abs(loss64.item()) < 1e-5
    (2.0 * loss).backward()
    assert _rel(rc.grad, 2.0 * r64.grad) < 1e-4
    assert _rel(dc.grad, 2.0 * d64.grad) < 2e-3          # fp32 cancellation at small dispersion (psi(y + 1/a) - psi(1/a))
    if with_mask:
        assert float(rc.grad[~mask.cuda()].abs().max()) == 0.0


def test_layer_norm_backward_matches_float64_autograd():
    from timesnet_forecast.autograd import layer_norm
    g = torch.Generator().manual_seed(1)
    x = torch.randn(6, 50, 96, generator=g)
    w = 1.0 + 0.1 * torch.randn(96, generator=g)
    b = 0.1 * torch.randn(96, generator=g)
    up = torch.randn(6, 50, 96, generator=g)
    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x, w, b))
    (F.layer_norm(x64, (96,), w64, b64, 1e-5) * up.double()).sum().backward()
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    out = layer_norm(xc, wc, bc, 1e-5)
    assert _rel(out, F.layer_norm(x64, (96,), w64, b64, 1e-5)) < 1e-5
    (out * up.cuda()).sum().backward()
    assert _rel(xc.grad, x64.grad) < 1e-4 and _rel(wc.grad, w64.grad) < 1e-4 and _rel(bc.grad, b64.grad) < 1e-4


@pytest.mark.parametrize("with_late", [False, True])
def test_nb_head_backward_matches_float64_autograd(with_late):
    from timesnet_forecast.autograd import nb_head
    g = torch.Generator().manual_seed(2)
    B, L, C, steps, N = 5, 40, 32, 12, 9
    seq = torch.randn(B, L, C, generator=g)
    Wt = torch.randn(steps, L, generator=g) / L ** 0.5
    bt = 0.1 * torch.randn(steps, generator=g)
    Wmu, Wsg = (0.2 * torch.randn(N, C, generator=g) for _ in range(2))
    bmu, bsg = (0.1 * torch.randn(N, generator=g) for _ in range(2))
    hist = torch.rand(B, steps, N, generator=g) * 3
    late = torch.randn(B, N, steps, generator=g) if with_late else None
    gate = (0.05 + 0.01 * torch.randn(steps, generator=g)) if with_late else None
    floor = 0.01 + 0.1 * torch.rand(N, generator=g)
    ur, ud = torch.randn(B, steps, N, generator=g), torch.randn(B, steps, N, generator=g)
    names = ["seq", "Wt", "bt", "Wmu", "bmu", "Wsg", "bsg"] + (["late", "gate"] if with_late else [])
    vals = dict(seq=seq, Wt=Wt, bt=bt, Wmu=Wmu, bmu=bmu, Wsg=Wsg, bsg=bsg, late=late, gate=gate)
    p64 = {k: vals[k].double().requires_grad_(True) for k in names}
    hidden = torch.einsum("ht,btc->bhc", p64["Wt"], p64["seq"]) + p64["bt"].view(1, -1, 1)
    pre = hidden @ p64["Wmu"].t() + p64["bmu"] + hist.double()
    if with_late:
        pre = pre + p64["gate"].view(1, -1, 1) * p64["late"].permute(0, 2, 1)
    rate64 = F.softplus(pre, beta=1.0, threshold=20) + 1e-6
    disp64 = F.softplus(hidden @ p64["Wsg"].t() + p64["bsg"], beta=1.0, threshold=20) + floor.double() + 1e-6
    ((rate64 * ur.double()).sum() + (disp64 * ud.double()).sum()).backward()
    pc = {k: vals[k].cuda().requires_grad_(True) for k in names}
    rate, disp = nb_head(pc["seq"], pc["Wt"], pc["bt"], pc["Wmu"], pc["bmu"], pc["Wsg"], pc["bsg"], hist.cuda(),
                         pc.get("late"), pc.get("gate"), floor.cuda())
    assert _rel(rate, rate64) < 1e-5 and _rel(disp, disp64) < 1e-5
    ((rate * ur.cuda()).sum() + (disp * ud.cuda()).sum()).backward()
    for k in names:
        assert _rel(pc[k].grad, p64[k].grad) < 1e-4, k


def test_bf16_blocks_stay_forward_only():
    """The differentiable route is fp32: a bf16 block refuses inputs that require grad instead of returning grad-less
    tensors silently."""
    from timesnet_forecast.models.timesnet import TimesBlock
    blk = TimesBlock(16, [(3, 3)], 0.0, "gelu").cuda().eval()

    class Sel(torch.nn.Module):
        def forward(self, x):
            return torch.tensor([4, 6]), torch.ones(x.size(0), 2, device=x.device, dtype=x.dtype)
    object.__setattr__(blk, "period_selector", Sel())
    with pytest.raises(RuntimeError):
        blk(torch.randn(2, 24, 16, device="cuda", dtype=torch.bfloat16, requires_grad=True))


@pytest.mark.parametrize("with_ln", [False, True])
def test_timesblock_backward_matches_float64_autograd(with_ln):
    """TimesBlock.forward on an fp32 input that requires grad: fold views + differentiable Inception chain + aggregation
    (all compute in libflowtimes, forward and backward) against a float64 torch evaluation of timesnet.py:767-818 with
    the same periods and the same (constant) group weights.  Periods 4, 6, 5: pads 0, 0 and 1 at L = 24."""
    import copy
    from timesnet_forecast.models.timesnet import TimesBlock
    torch.manual_seed(5)
    C, L, B = 16, 24, 2
    blk = TimesBlock(C, [(3, 3), (5, 5)], 0.0, "gelu", d_ff=32, bottleneck_ratio=4.0)
    for p in blk.parameters():
        p.data.normal_(0, 0.25)
    ref = copy.deepcopy(blk.inception).double()
    amps = torch.tensor([[1.0, 0.5, 0.2], [0.3, 0.9, 0.1]])

    class Sel(torch.nn.Module):
        def forward(self, x):
            return torch.tensor([4, 6, 5]), amps.to(device=x.device, dtype=x.dtype)
    blk = blk.cuda().eval()
    object.__setattr__(blk, "period_selector", Sel())
    norm = torch.nn.LayerNorm(C)
    norm.weight.data.normal_(1, 0.1)
    norm.bias.data.normal_(0, 0.1)
    x = torch.randn(B, L, C)
    u = torch.randn(B, L, C)

    def inc(m, g):
        out = m.act(m.proj(torch.cat([p.branch(g) for p in m.paths], dim=1)))
        return out + m.res_proj(g)
    xr = x.double().requires_grad_()
    w = torch.softmax(amps.double(), dim=1)
    acc = xr
    for gi, p in enumerate([4, 5, 6]):                                # groups ascend by period
        col = {4: 0, 5: 2, 6: 1}[p]
        pad = (-L) % p
        grid = torch.nn.functional.pad(xr, (0, 0, 0, pad)).reshape(B, (L + pad) // p, p, C).permute(0, 3, 1, 2)
        y = inc(ref[2], ref[1](inc(ref[0], grid)))
        d = (y - grid).permute(0, 2, 3, 1).reshape(B, L + pad, C)[:, :L]
        acc = acc + w[:, col].view(B, 1, 1) * d
    n64 = copy.deepcopy(norm).double()
    want = n64(acc) if with_ln else acc
    (want * u.double()).sum().backward()
    xc = x.cuda().requires_grad_()
    normc = norm.cuda()
    got = blk.forward_norm_differentiable(xc, normc) if with_ln else blk(xc)
    assert got.requires_grad and _rel(got, want) < 2e-5
    (got * u.cuda()).sum().backward()
    assert _rel(xc.grad, xr.grad) < 5e-5
    for (n, p), (_, q) in zip(blk.inception.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and _rel(p.grad, q.grad) < 5e-5, n
    if with_ln:
        assert _rel(normc.weight.grad, n64.weight.grad) < 5e-5 and _rel(normc.bias.grad, n64.bias.grad) < 5e-5


# ---- second slice: Inception chain pieces and the aggregation -------------------------------------------------------
@pytest.mark.parametrize("kh,kw", [(1, 1), (3, 3), (5, 7)])
def test_conv2d_same_backward_matches_float64_autograd(kh, kw):
    """conv2d with zero "same" padding on an NCHW grid: native forward (ftn_conv2d_grid on the fold), native data /
    weight / bias gradients, against float64 F.conv2d autograd (InceptionBranch's convs, timesnet.py:575-593)."""
    from timesnet_forecast.autograd import conv2d_same
    g = torch.Generator().manual_seed(kh * 10 + kw)
    B, Cin, Cout, H, W = 2, 5, 7, 4, 6
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5
    b = torch.randn(Cout, generator=g)
    u = torch.randn(B, Cout, H, W, generator=g)
    ref = [t.double().requires_grad_() for t in (x, w, b)]
    (torch.nn.functional.conv2d(ref[0], ref[1], ref[2], padding=(kh // 2, kw // 2)) * u.double()).sum().backward()
    got = [t.cuda().requires_grad_() for t in (x, w, b)]
    out = conv2d_same(got[0], got[1], got[2])
    want = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=(kh // 2, kw // 2))
    assert _rel(out, want) < 1e-5
    (out * u.cuda()).sum().backward()
    for name, a, r in zip(("dx", "dw", "db"), got, ref):
        assert _rel(a.grad, r.grad) < 1e-5, name


@pytest.mark.parametrize("name", ["gelu", "relu"])
def test_activation_backward_matches_float64_autograd(name):
    from timesnet_forecast.autograd import activation
    x = torch.linspace(-6, 6, 4001)
    r = x.double().requires_grad_()
    (torch.nn.functional.gelu(r) if name == "gelu" else torch.relu(r)).sum().backward()
    c = x.cuda().requires_grad_()
    activation(c, name).sum().backward()
    assert (c.grad.cpu().double() - r.grad).abs().max().item() < 2e-6


@pytest.mark.parametrize("ratio,cin,cout", [(4.0, 8, 16), (1.0, 6, 6), (2.0, 12, 8)])
def test_inception_block_backward_matches_float64_autograd(ratio, cin, cout):
    """InceptionBlock.forward on an NCHW grid that requires grad takes the differentiable route: outputs and every
    gradient (input, all branch convs, proj, res_proj) against the same module evaluated in float64 with torch ops
    (the reference's formula, timesnet.py:645-654)."""
    import copy
    from timesnet_forecast.models.timesnet import InceptionBlock
    torch.manual_seed(int(ratio * 10) + cin)
    blk = InceptionBlock(cin, cout, [(3, 3), (5, 5)], 0.0, "gelu", bottleneck_ratio=ratio)
    for p in blk.parameters():
        p.data.normal_(0, 0.3)
    ref = copy.deepcopy(blk).double()
    x = torch.randn(2, cin, 3, 8)
    u = torch.randn(2, cout, 3, 8)

    def ref_forward(m, xx):
        feats = [p.branch(xx) for p in m.paths]
        out = m.act(m.proj(torch.cat(feats, dim=1)))
        return out + m.res_proj(xx)
    xr = x.double().requires_grad_()
    want = ref_forward(ref, xr)
    (want * u.double()).sum().backward()
    blk = blk.cuda()
    xc = x.cuda().requires_grad_()
    got = blk(xc)
    assert got.requires_grad and _rel(got, want) < 1e-5
    (got * u.cuda()).sum().backward()
    assert _rel(xc.grad, xr.grad) < 1e-5
    for (n, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        assert p.grad is not None and _rel(p.grad, q.grad) < 1e-5, n
    # a grad-less input keeps the fast forward-only route (packed chain), bit-for-bit deterministic
    with torch.no_grad():
        assert not blk(x.cuda()).requires_grad


def test_aggregate_backward_matches_float64_autograd():
    from timesnet_forecast import _native as nv
    from timesnet_forecast.autograd import aggregate
    g = torch.Generator().manual_seed(3)
    B, L, C = 3, 24, 16
    plan_host = nv.plan_build_host([4, 6, 12, 6], L, None, None)       # 3 groups (6 twice)
    plan = nv.plan_to_device(plan_host, "cuda")
    G = plan_host.n_groups
    x = torch.randn(B, L, C, generator=g)
    delta = torch.randn(nv.FTN_MAX_K, B, L, C, generator=g)
    w = torch.zeros(B, nv.FTN_MAX_K)
    w[:, :G] = torch.softmax(torch.randn(B, G, generator=g), dim=1)
    u = torch.randn(B, L, C, generator=g)
    xr, dr, wr = x.double().requires_grad_(), delta.double().requires_grad_(), w.double().requires_grad_()
    want = xr + sum(wr[:, gi].view(B, 1, 1) * dr[gi] for gi in range(G))
    (want * u.double()).sum().backward()
    xc, dc, wc = x.cuda().requires_grad_(), delta.cuda().requires_grad_(), w.cuda().requires_grad_()
    got = aggregate(xc, dc, wc, plan)
    assert _rel(got, want) < 1e-6
    (got * u.cuda()).sum().backward()
    assert _rel(xc.grad, xr.grad) < 1e-6
    assert _rel(dc.grad[:G], dr.grad[:G]) < 1e-6 and float(dc.grad[G:].abs().max()) == 0.0
    assert _rel(wc.grad[:, :G], wr.grad[:, :G]) < 1e-5


def test_timesblock_backward_through_fft_selector():
    """With the FFT selector the gradient also flows through the group weights (softmax <- amplitudes <- |rfft| of the
    median channel), like the reference's autograd (timesnet.py:992-1009; tests/test_fft_period_selector.py:73-102
    pins that amplitudes carry grad).  Float64 torch evaluation with the SAME bins (top-k has no gradient)."""
    import copy
    from timesnet_forecast.models.timesnet import FFTPeriodSelector, TimesBlock
    torch.manual_seed(11)
    C, L, B, k = 8, 36, 3, 3
    blk = TimesBlock(C, [(3, 3)], 0.0, "gelu", d_ff=16, bottleneck_ratio=2.0)
    for p in blk.parameters():
        p.data.normal_(0, 0.3)
    ref = copy.deepcopy(blk.inception).double()
    blk = blk.cuda().eval()
    object.__setattr__(blk, "period_selector", FFTPeriodSelector(k, L, 1))
    t = torch.arange(L, dtype=torch.float32).view(1, L, 1)
    x = (torch.sin(2 * torch.pi * t / 6 + torch.rand(B, 1, C) * 6) + 0.7 * torch.sin(2 * torch.pi * t / 9 + torch.rand(B, 1, C))
         + 0.5 * torch.sin(2 * torch.pi * t / 4) + 0.3 * torch.randn(B, L, C))
    u = torch.randn(B, L, C)
    xc = x.cuda().requires_grad_()
    got = blk(xc)
    h = blk._last_plan.host()
    nvld, G = h.n_valid, h.n_groups
    bins = [int(h.freq[j]) for j in range(nvld)]
    mapping = [int(h.mapping[j]) for j in range(nvld)]
    assert G >= 2

    def inc(m, g):
        out = m.act(m.proj(torch.cat([p.branch(g) for p in m.paths], dim=1)))
        return out + m.res_proj(g)
    xr = x.double().requires_grad_()
    amp = torch.fft.rfft(xr, dim=1).abs().median(dim=2).values                      # [B, F]
    a = amp[:, bins]
    valid = [j for j in range(nvld) if mapping[j] >= 0]
    sm = torch.softmax(a[:, valid], dim=1)
    w = [sum(sm[:, i] for i, j in enumerate(valid) if mapping[j] == g) for g in range(G)]
    acc = xr
    for g in range(G):
        p, pad, cyc = int(h.grp_period[g]), int(h.grp_pad[g]), int(h.grp_cycles[g])
        grid = torch.nn.functional.pad(xr, (0, 0, 0, pad)).reshape(B, cyc, p, C).permute(0, 3, 1, 2)
        y = inc(ref[2], ref[1](inc(ref[0], grid)))
        acc = acc + w[g].view(B, 1, 1) * (y - grid).permute(0, 2, 3, 1).reshape(B, cyc * p, C)[:, :L]
    (acc * u.double()).sum().backward()
    assert _rel(got, acc) < 2e-5
    (got * u.cuda()).sum().backward()
    assert _rel(xc.grad, xr.grad) < 1e-4
    # the weight path is really there: with constant weights the input gradient differs measurably
    blk2_in = x.cuda().requires_grad_()
    from timesnet_forecast import autograd as ag
    keep = ag.period_weights
    try:
        ag.period_weights = lambda xx, plan: plan.weights.detach()
        (blk(blk2_in) * u.cuda()).sum().backward()
    finally:
        ag.period_weights = keep
    assert _rel(blk2_in.grad, xr.grad) > 1e-3


@pytest.mark.parametrize("name", ["toy_direct", "toy_recursive"])
def test_timesnet_training_step_gradients_match_oracle_autograd(golden_dir, name):
    """One training step's gradients through the WHOLE model (embedding, context, TimesBlock stack with the period-weight
    path, NB head, NB-NLL) on the differentiable route, against torch autograd of the oracle (= the reference's forward,
    pinned by the goldens) on the CPU with the same weights.  fp32 on both sides: 2e-3 of the largest gradient."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    import flowtimes_oracle as orc
    import flowtimes_synth as syn
    from test_gpu_parity import _build_model, _wl
    from timesnet_forecast.autograd import nb_nll
    c = torch.load(golden_dir / "model_toy.pt")[name]
    wl = _wl(c["workload"])
    m, x, static, ids = _build_model(c, wl)
    m.dropout = 0.0
    for b in m.blocks:
        b._dropout = 0.0
    m.differentiable = True
    for b in m.blocks:
        b.differentiable = True
    y, mask = c["y"].cuda(), c["mask"].cuda()
    rate, disp = m(x, series_static=static, series_ids=ids)
    assert rate.requires_grad
    steps = wl.H if wl.mode == "direct" else 1
    loss = nb_nll(y[:, :steps], rate, disp, mask[:, :steps].to(torch.uint8).contiguous() if mask is not None else None)
    loss.backward()
    # oracle autograd
    w = {k: v.clone().float().requires_grad_(v.is_floating_point()) for k, v in c["state"].items()}
    cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu", wl.min_period_threshold, 1e-3,
                       wl.context_rank > 0, wl.context_rank)
    xr = syn.planted_series(wl.B, c["T_in"], wl.N, seed=c["x_seed"])
    r, d = orc.timesnet_forward(xr, w, cfg, series_static=c["static"], series_ids=c["ids"],
                                min_sigma_vector=c["min_sigma_vector"])
    lo = orc.nb_nll(c["y"][:, :steps], r, d, c["mask"][:, :steps] if c["mask"] is not None else None)
    lo.backward()
    assert _rel(loss, lo) < 1e-4
    checked = 0
    for k, p in m.named_parameters():
        if w[k].grad is None:
            continue
        assert p.grad is not None, f"{k}: no gradient on the differentiable route"
        gr = w[k].grad
        if float(gr.abs().max()) < 1e-10:
            assert float(p.grad.abs().max()) < 1e-6, k
            continue
        assert _rel(p.grad, gr) < 2e-3, f"{k}: {_rel(p.grad, gr):.2e}"
        checked += 1
    assert checked >= 20
