"""GPU: unit tests of the tcgen05 / TMEM / TMA GEMM kernel in isolation (bf16 in, fp32 accumulate,
bf16 out) against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (256, 96, 96), (128, 512, 128), (384, 128, 512), (1280, 48, 64),
                                   (128, 16, 16)])
def test_tc_linear_matches_fp32_matmul(M, K, N):
    from timesnet_forecast import _native as nv
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    out = nv.debug_tc_linear(a.cuda(), w.cuda(), bias.cuda())
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    err = (out.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 8e-3, f"rel err {err:.3e}"          # one bf16 rounding of the output
    # structured input: catches swizzle / descriptor mistakes that random data could hide
    a2 = torch.zeros(M, K)
    a2[torch.arange(M), torch.arange(M) % K] = 1.0
    out2 = nv.debug_tc_linear(a2.to(torch.bfloat16).cuda(), w.cuda(), torch.zeros(N).cuda())
    ref2 = w.float().t()[torch.arange(M) % K]
    assert torch.equal(out2.float().cpu(), ref2)
