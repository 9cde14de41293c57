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
