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


def test_blocks_stay_forward_only():
    """The Inception chain has no backward yet: a block refuses inputs that require grad instead of returning
    grad-less tensors silently."""
    from timesnet_forecast.models.timesnet import TimesBlock
    blk = TimesBlock(16, [(3, 3)], 0.0, "gelu").cuda().eval()

    class Sel(torch.nn.Module):
        def forward(self, x):
            return torch.tensor([4, 6]), torch.ones(x.size(0), 2, device=x.device, dtype=x.dtype)
    object.__setattr__(blk, "period_selector", Sel())
    with pytest.raises(RuntimeError):
        blk(torch.randn(2, 24, 16, device="cuda", requires_grad=True))
