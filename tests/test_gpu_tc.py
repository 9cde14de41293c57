"""GPU: unit tests of the tcgen05 / TMEM / TMA GEMM kernel in isolation (bf16 in, fp32 accumulate,
bf16 out) against a plain PyTorch fp32 reference of the same op."""
import math

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


@pytest.mark.parametrize("mid,L,periods", [(32, 336, [24, 12, 7, 48, 6]), (32, 336, [335, 100, 168]), (32, 336, [2, 3, 5]),
                                            (16, 96, [24, 12, 7, 48, 6, 95]), (32, 28, [27, 14, 7]), (32, 96, [1, 2, 48]),
                                            (32, 720, [6, 24, 359])])
@pytest.mark.parametrize("variant", [2, 4])
def test_tc_conv_matches_simt_conv(mid, L, periods, variant):
    if variant == 4 and mid != 32:
        pytest.skip("tc_conv4 is written for mid = 32")
    """k x k stage alone: tcgen05 implicit-GEMM kernel vs the fp32-math SIMT kernel on the same
    tile-major bf16 activations (identical inputs, fp32 accumulation in both -> <= 1 bf16 ulp)."""
    _conv_vs_simt(mid, L, periods, variant, B=3)


@pytest.mark.parametrize("mid,L,periods,B", [(16, 96, [24, 12, 7, 48, 6], 256), (32, 96, [24, 12, 7, 48, 6], 128),
                                              (16, 336, [24, 12, 7, 48, 335], 40)])
def test_tc_conv2_many_units_per_cta(mid, L, periods, B):
    """Image-resident kernel with tens of units per persistent CTA (the etth1 batch: 256 windows x 5 groups over 148
    CTAs).  mid = 16 splits the epilogue warps by tile parity; on single-tile images the odd-tile warps own nothing
    and used to run ahead of the unit, breaking the accumulator hand-over (trap after a bounded spin)."""
    _conv_vs_simt(mid, L, periods, 2, B=B)


def _conv_vs_simt(mid, L, periods, variant, B):
    import flowtimes_synth as syn  # noqa: F401
    from timesnet_forecast import _native as nv
    from timesnet_forecast.models.timesnet import InceptionBlock
    C = mid * 4
    torch.manual_seed(0)
    blk = InceptionBlock(C, C, [(3, 3), (5, 5), (7, 7)], 0.0, "gelu", bottleneck_ratio=4.0).cuda()
    packed = blk.packed(torch.device("cuda"))
    assert packed.struct.mid == mid
    plan_host = nv.plan_build_host(periods, L, None, None)
    plan = nv.plan_to_device(plan_host, "cuda")
    G = plan_host.n_groups
    tiles = len(periods) * B * ((2 * L + 127) // 128)
    NB = 3 * mid
    g = torch.Generator().manual_seed(1)
    inp = torch.randn(tiles * 128, NB, generator=g).to(torch.bfloat16).cuda()
    ref = nv.debug_conv_tiled(inp, plan, B, L, len(periods), packed.struct, use_tc=False)
    got = nv.debug_conv_tiled(inp, plan, B, L, len(periods), packed.struct, use_tc=variant)
    torch.cuda.synchronize()
    # compare only rows that belong to an image (pad rows of a tile are never written)
    row = 0
    checked = 0
    for gi in range(G):
        Lp = L + plan_host.grp_pad[gi]
        rt = (Lp + 127) // 128
        for b in range(B):
            a = ref[row:row + Lp].float()
            c = got[row:row + Lp].float()
            err = (a - c).abs().max().item() / max(1e-6, a.abs().max().item())
            assert err < 1e-2, f"group {gi} (p={plan_host.grp_period[gi]}) window {b}: rel err {err:.3e}"
            checked += Lp
            row += rt * 128
    assert checked > 0


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (256, 48, 64), (128, 256, 48), (384, 128, 256), (128, 16, 16), (256, 1024, 192)])
def test_tc_linear_split_keeps_fp32_accuracy(M, K, N):
    """Three-plane bf16 GEMM (6 MMAs per product, fp32 accumulate) against a float64 matmul: fp32-level error, four
    orders of magnitude below a single bf16 product."""
    from timesnet_forecast import _native as nv
    from timesnet_forecast._pack import split3
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    w_s3 = torch.cat(split3(w), dim=1).contiguous()
    out = nv.debug_tc_linear_split(a.cuda(), w_s3.cuda(), bias.cuda())
    torch.cuda.synchronize()
    ref = (a.double() @ w.double().t() + bias.double())
    err = (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, f"rel err {err:.3e}"


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (256, 48, 64), (128, 256, 48), (384, 128, 256), (128, 16, 16), (256, 1024, 192)])
@pytest.mark.parametrize("wscale", [1.0, 1e-3, 300.0])
def test_tc_linear_h2_keeps_fp32_accuracy(M, K, N, wscale):
    """Two-plane fp16 GEMM (3 MMAs per product, fp32 accumulate, power-of-two weight scaling) against a float64 matmul:
    fp32-level error whatever the magnitude of the weights."""
    from timesnet_forecast import _native as nv
    from timesnet_forecast._pack import split_h2
    g = torch.Generator().manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g) * 3.0
    a[0, 0] = 1e-4                                     # deep in the range where the lo plane is an fp16 subnormal
    w = torch.randn(N, K, generator=g) / K ** 0.5 * wscale
    bias = torch.randn(N, generator=g) * wscale
    planes, inv = split_h2(w)
    w_h2 = torch.cat(planes, dim=1).contiguous()
    out = nv.debug_tc_linear_h2(a.cuda(), w_h2.cuda(), inv, bias.cuda())
    torch.cuda.synchronize()
    ref = (a.double() @ w.double().t() + bias.double())
    err = (out.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, f"rel err {err:.3e}"


def test_split_h2_saturates_and_keeps_nan():
    """fp16 planes cannot hold |v| > 65504: finite values saturate, NaN propagates (documented limit of the h2 form)."""
    from timesnet_forecast import _native as nv
    from timesnet_forecast._pack import split_h2
    a = torch.zeros(128, 16)
    a[0, 0] = 1e6
    a[1, 0] = float("nan")
    w = torch.eye(16)
    planes, inv = split_h2(w)
    out = nv.debug_tc_linear_h2(a.cuda(), torch.cat(planes, dim=1).contiguous().cuda(), inv, torch.zeros(16).cuda()).cpu()
    assert out[0, 0].item() == 65504.0
    assert math.isnan(out[1, 0].item())
    assert torch.equal(out[2:], torch.zeros(126, 16))


def _tile_major_images(plan_host, B, L, NB, seed, dtype):
    """Random per-image activations placed in the tile-major layout of the tensor-core chain."""
    G = plan_host.n_groups
    tiles = sum(B * ((L + plan_host.grp_pad[g] + 127) // 128) for g in range(G))
    g_ = torch.Generator().manual_seed(seed)
    buf = torch.zeros(tiles * 128, NB, dtype=dtype)
    imgs, row = [], 0
    for gi in range(G):
        Lp = L + plan_host.grp_pad[gi]
        rt = (Lp + 127) // 128
        for b in range(B):
            im = torch.randn(Lp, NB, generator=g_).to(dtype)
            buf[row:row + Lp] = im
            imgs.append((gi, row, im))
            row += rt * 128
    return buf, imgs


@pytest.mark.parametrize("mid,L,periods", [(16, 96, [24, 12, 7, 48, 95]), (32, 336, [24, 168, 335, 5]), (64, 720, [6, 24, 359, 719]),
                                            (64, 96, [1, 2, 48]), (48, 50, [7, 25]),
                                            # mid = 16 runs the row mode (taps of a tap row on N): long periods (one segment
                                            # per tap row), single-row grids, periods shorter than the kernel
                                            (16, 336, [24, 12, 7, 168, 335]), (16, 720, [6, 359, 719]), (16, 50, [1, 2, 3, 49])])
@pytest.mark.parametrize("planes", [1, 2, 3])
def test_tc_convs_streaming_kernel_matches_conv2d(mid, L, periods, planes):
    """Streaming k x k kernel (any mid, bf16 or three-plane fp32 activations) against torch conv2d in float64 on the
    folded grids: bf16 activations to one output rounding, the three-plane mode to fp32 accuracy."""
    _convs_vs_conv2d(mid, L, periods, planes, [(3, 3), (5, 5), (7, 7)])


@pytest.mark.parametrize("mid", [16, 64])
@pytest.mark.parametrize("planes", [1, 2])
def test_tc_convs_rectangular_kernels(mid, planes):
    """Non-square kernels (kh != kw, a 1 x 7 row kernel, a 7 x 1 column kernel) through the streaming kernel: mid = 16
    takes the row mode (kw taps on N, kh MMAs), mid = 64 the tap-by-tap stream."""
    _convs_vs_conv2d(mid, 96, [24, 7, 48, 95], planes, [(3, 5), (1, 7), (7, 1)])


def _convs_vs_conv2d(mid, L, periods, planes, kernel_set):
    from timesnet_forecast import _native as nv
    from timesnet_forecast._pack import split3, split_h2
    from timesnet_forecast.models.timesnet import InceptionBlock
    B, C = 2, mid * 4
    torch.manual_seed(0)
    blk = InceptionBlock(C, C, kernel_set, 0.0, "gelu", bottleneck_ratio=4.0).cuda()
    packed = blk.packed(torch.device("cuda"))
    assert packed.struct.mid == mid
    plan_host = nv.plan_build_host(periods, L, None, None)
    plan = nv.plan_to_device(plan_host, "cuda")
    NB = 3 * mid
    dt = torch.bfloat16 if planes == 1 else torch.float32
    buf, imgs = _tile_major_images(plan_host, B, L, NB, 1, dt)
    if planes == 1:
        inp = buf.cuda()
    elif planes == 2:
        hi = buf.to(torch.float16)
        lo = (buf - hi.float()).to(torch.float16)
        inp = torch.cat([hi, lo], dim=1).contiguous().cuda()           # [rows][2 NB] fp16 (activations are not scaled)
    else:
        inp = torch.cat(split3(buf), dim=1).contiguous().cuda()        # [rows][3 NB]
    got = nv.debug_conv_tiled(inp, plan, B, L, len(periods), packed.struct, use_tc={1: 5, 2: 7, 3: 6}[planes])
    torch.cuda.synchronize()
    got = got.float().cpu()
    if planes == 3:
        got = got[:, :NB] + got[:, NB:2 * NB] + got[:, 2 * NB:]
    elif planes == 2:
        got = got[:, NB:] + got[:, :NB]
    worst = 0.0
    for gi, row, im in imgs:
        per, cyc = plan_host.grp_period[gi], plan_host.grp_cycles[gi]
        Lp = per * cyc
        for j, path in enumerate(blk.paths):
            conv = path.branch[1]
            grid = im[:, j * mid:(j + 1) * mid].double().t().reshape(1, mid, cyc, per)
            ref = torch.nn.functional.conv2d(grid, conv.weight.detach().double().cpu(), conv.bias.detach().double().cpu(),
                                             padding=(conv.kernel_size[0] // 2, conv.kernel_size[1] // 2))
            ref = ref.reshape(mid, Lp).t()
            out = got[row:row + Lp, j * mid:(j + 1) * mid].double()
            err = (out - ref).abs().max().item() / ref.abs().max().item()
            worst = max(worst, err)
            assert err < (1e-2 if planes == 1 else 3e-5), f"group {gi} (p={per}) branch {j}: rel err {err:.3e}"
    assert worst > 0.0
