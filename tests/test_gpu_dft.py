"""Tensor-core spectrum (csrc/tc_dft.cu): the DFT-as-GEMM route of the period search against the SIMT FFT route of the
same library and against the oracle's channel-median spectrum (timesnet.py:109-111)."""
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
sys.path.insert(0, str(ROOT))


def _oracle_median(x):
    xf = x.float().cpu()
    amp = torch.fft.rfft(xf, dim=1).abs()
    return amp.median(dim=2).values          # lower median, NaN-propagating (torch.median)


@pytest.mark.parametrize("B,L,C", [(8, 336, 128), (5, 96, 64), (3, 720, 128), (2, 130, 64), (64, 336, 128)])
def test_tensor_dft_matches_fft_route_and_oracle(B, L, C):
    from timesnet_forecast import _native as nv
    g = torch.Generator().manual_seed(1000 + L + C)
    t = torch.arange(L, dtype=torch.float32).view(1, L, 1)
    x = (5.0 + 3.0 * torch.sin(2 * torch.pi * t / 24 + torch.rand(B, 1, C, generator=g) * 6.28)
         + torch.randn(B, L, C, generator=g)).to(torch.bfloat16).cuda()
    assert nv.dft_basis(x) is not None, "shape should take the tensor-core route"
    k = 5
    plan_t, amps_t, w_t, med_t, sum_t = nv.period_search(x, k, L, 1)
    plan_s, amps_s, w_s, med_s, sum_s = nv.period_search(x, k, L, 1, tensor_dft=False)
    torch.cuda.synchronize()
    ref = _oracle_median(x)
    scale = ref.abs().max().item()
    err_t = (med_t.cpu() - ref).abs().max().item() / scale
    err_s = (med_s.cpu() - ref).abs().max().item() / scale
    assert err_t <= 1e-5, f"tensor-core spectrum off by {err_t:.2e} (SIMT FFT route: {err_s:.2e})"
    assert torch.equal(plan_t.cpu()[:804], plan_s.cpu()[:804]), "the two routes select different periods"
    assert (amps_t.float() - amps_s.float()).abs().max().item() <= 2e-2 * max(1.0, amps_s.float().abs().max().item())


def test_tensor_dft_propagates_nan_like_torch_median():
    from timesnet_forecast import _native as nv
    B, L, C = 4, 336, 128
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, L, C, generator=g).to(torch.bfloat16)
    x[1, 17, 5] = float("nan")
    x = x.cuda()
    _, _, _, med, _ = nv.period_search(x, 3, L, 1)
    med = med.cpu()
    assert torch.isnan(med[1]).all(), "a NaN sample poisons every bin of its window (torch.median propagates NaN)"
    assert torch.isfinite(med[[0, 2, 3]]).all()
    ref = _oracle_median(x)
    assert torch.allclose(med[[0, 2, 3]], ref[[0, 2, 3]], rtol=0, atol=1e-5 * ref[[0, 2, 3]].abs().max().item())


def test_tensor_dft_off_switch_and_ineligible_shapes():
    from timesnet_forecast import _native as nv
    assert nv.dft_basis(torch.zeros(2, 336, 128, device="cuda")) is None                            # fp32 activations
    assert nv.dft_basis(torch.zeros(2, 336, 96, device="cuda", dtype=torch.bfloat16)) is None       # C not 64 / 128
    assert nv.dft_basis(torch.zeros(2, 28, 128, device="cuda", dtype=torch.bfloat16)) is None       # short window


@pytest.mark.parametrize("B,L,C,steps,N", [(8, 336, 128, 96, 321), (6, 96, 64, 96, 32), (3, 720, 256, 336, 40),
                                            (64, 336, 128, 96, 321), (5, 100, 128, 7, 16)])
def test_time_projection_on_tensor_cores_matches_simt_and_fp64(B, L, C, steps, N):
    """forecast_time_proj (timesnet.py:2071) through the MN-major tcgen05 GEMM (tc_dft.cu MODE 1) against the fp32 SIMT
    projection of the same head call and against a float64 evaluation of the whole head."""
    from timesnet_forecast import _native as nv
    from timesnet_forecast._pack import split_linear_weight
    g = torch.Generator().manual_seed(L * 7 + C)
    seq = torch.randn(B, L, C, generator=g).to(torch.bfloat16).cuda()
    Wt = (torch.randn(steps, L, generator=g) / L ** 0.5).cuda()
    bt = (0.1 * torch.randn(steps, generator=g)).cuda()
    Wmu, Wsg = (torch.randn(N, C, generator=g) / C ** 0.5).cuda(), (torch.randn(N, C, generator=g) / C ** 0.5).cuda()
    bmu, bsg = (0.1 * torch.randn(N, generator=g)).cuda(), (0.1 * torch.randn(N, generator=g)).cuda()
    n_pad = (N + 127) // 128 * 128
    w = split_linear_weight(Wmu, C, 2 * n_pad, 0) + split_linear_weight(Wsg, C, 2 * n_pad, n_pad)
    b = torch.zeros(2 * n_pad, device="cuda")
    b[:N], b[n_pad:n_pad + N] = bmu, bsg
    hist = torch.rand(B, steps, N, generator=g).cuda()
    floor = torch.full((N,), 1e-3, device="cuda")
    out = {}
    for route in ("tensor", "simt"):
        flags = torch.zeros(1, dtype=torch.int32, device="cuda")
        wt_s3 = nv.time_proj_pack(Wt) if route == "tensor" else None
        res = nv.nb_head_tc(seq, steps, N, Wt, bt, w.contiguous(), b, n_pad, hist, None, None, floor, flags, wt_s3=wt_s3)
        assert res is not None
        out[route] = [t.cpu() for t in res]
        assert int(flags.item()) == 0
    hidden = torch.einsum("ht,btc->bhc", Wt.double().cpu(), seq.double().cpu()) + bt.double().cpu()[None, :, None]
    rate = torch.nn.functional.softplus(hidden @ Wmu.double().cpu().T + bmu.double().cpu() + hist.double().cpu()) + 1e-6
    disp = torch.nn.functional.softplus(hidden @ Wsg.double().cpu().T + bsg.double().cpu()) + 1e-3 + 1e-6
    for name, ref, i in (("rate", rate, 0), ("dispersion", disp, 1)):
        scale = ref.abs().max().item()
        for route in ("tensor", "simt"):
            err = (out[route][i].double() - ref).abs().max().item() / scale
            assert err <= 1e-5, f"{name} via the {route} projection off by {err:.2e}"
