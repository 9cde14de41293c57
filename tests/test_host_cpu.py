"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the host-resident integer logic (plan builder, PeriodGrouper, weight packing) agrees with
the oracle, the product path refuses CPU tensors (no fallback), and the N>1 protocol (shard the
batch, all-reduce the batch-summed spectrum, select identically on every rank) is exercised with
two ``gloo`` processes."""
import os
import re
import socket
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _header_symbols():
    text = (ROOT / "include" / "flowtimes.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"FTN_API\s+[\w\s\*]+?\b(ftn_\w+)\s*\(", text)))


def test_abi_exports_every_declared_symbol():
    import ctypes
    from timesnet_forecast import _native as nv
    syms = _header_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(str(nv.LIB_PATH))
    for s in syms:
        assert hasattr(lib, s), f"libflowtimes.so does not export {s}"
    assert set(syms) == set(nv.SIGNATURES), "ctypes signature table and include/flowtimes.h drifted apart"
    assert nv.load().ftn_version() == nv.ABI_VERSION
    assert nv.PLAN_BYTES == ctypes.sizeof(nv.FtnPeriodPlan)


def test_library_has_sm100a_code_only():
    import subprocess
    from timesnet_forecast import _native as nv
    out = subprocess.run(["cuobjdump", "-lelf", str(nv.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("L,periods", [(336, [24, 12, 7, 48, 6]), (96, [24, 12, 24, 8, 12]), (28, [27, 14, 28, 0]),
                                        (48, [47, 47, 2, 1, 100])])
def test_host_plan_builder_matches_oracle_grouping(L, periods):
    import flowtimes_oracle as orc
    from timesnet_forecast import _native as nv
    plan = nv.plan_build_host(periods, L, None, None)
    amps = torch.ones(2, len(periods))
    g = orc.group_periods(torch.tensor(periods), amps, L)
    G = plan.n_groups
    assert list(plan.grp_period[:G]) == [int(p) for p in g.periods]
    assert list(plan.grp_pad[:G]) == [int(p) for p in g.pads]
    assert list(plan.grp_cycles[:G]) == [int(p) for p in g.cycles]
    assert [m for m in plan.mapping[:len(periods)]] == [int(m) for m in g.mapping]
    off = 0
    for i in range(G):
        assert plan.grp_row_off[i] == off
        off += L + plan.grp_pad[i]
    assert plan.total_rows_per_window == off


def test_period_grouper_class_matches_oracle():
    import flowtimes_oracle as orc
    from timesnet_forecast.models.timesnet import PeriodGrouper
    torch.manual_seed(0)
    periods = torch.tensor([8, 4, 4, 3, 3, 50])
    amps = torch.rand(3, 6) * 4
    res = PeriodGrouper(periods, amps, 48, min_period=1, max_period=47, block_index=0, freq_indices=None).group()
    g = orc.group_periods(periods, amps, 48, min_period=1, max_period=47)
    assert res.periods.tolist() == [int(p) for p in g.periods]
    assert res.pad_lengths.tolist() == [int(p) for p in g.pads]
    assert res.cycles.tolist() == [int(p) for p in g.cycles]
    assert res.mapping.tolist() == [int(m) for m in g.mapping]
    assert torch.allclose(res.logits, g.logits, atol=1e-6)


@pytest.mark.parametrize("ratio,cin,cout", [(4.0, 16, 32), (4.0, 32, 16), (1.0, 8, 8), (2.0, 16, 16)])
def test_weight_packing_is_exact_refactoring(ratio, cin, cout):
    """proj o branch-out folding (and the ratio-1 union-window fold) reproduces the unfolded block."""
    import torch.nn.functional as F
    import flowtimes_oracle as orc
    from timesnet_forecast._pack import pack_inception_block
    from timesnet_forecast.models.timesnet import InceptionBlock
    import ctypes as C
    torch.manual_seed(1)
    blk = InceptionBlock(cin, cout, [(3, 3), (5, 5), (7, 7)], 0.0, "gelu", bottleneck_ratio=ratio)
    w = {"b." + k: v.detach() for k, v in blk.state_dict().items()}
    x = torch.randn(2, cin, 6, 9)
    ref = orc.inception_block(x, w, "b.", "gelu")
    pk = pack_inception_block(blk, torch.device("cpu"))
    st = pk.struct

    def arr(ptr, *shape):
        n = 1
        for s in shape:
            n *= s
        buf = (C.c_float * n).from_address(ptr)
        return torch.frombuffer(buf, dtype=torch.float32).clone().reshape(*shape)

    nb = st.n_branch
    if st.mid > 0:
        mid = st.mid
        h = torch.einsum("bchw,cn->bnhw", x, arr(st.w_in, cin, nb * mid)) + arr(st.b_in, nb * mid).view(1, -1, 1, 1)
        outs = []
        for j in range(nb):
            kh, kw = st.kh[j], st.kw[j]
            wk = arr(st.w_kk[j], kh * kw, mid, mid).reshape(kh, kw, mid, mid).permute(3, 2, 0, 1)   # OIHW
            outs.append(F.conv2d(h[:, j * mid:(j + 1) * mid], wk, arr(st.b_kk[j], mid), padding=(kh // 2, kw // 2)))
        h2 = torch.cat(outs, 1)
        z = torch.einsum("bchw,cn->bnhw", h2, arr(st.w_out, nb * mid, cout)) + arr(st.b_out, cout).view(1, -1, 1, 1)
    else:
        kh, kw = st.kh[0], st.kw[0]
        wk = arr(st.w_kk[0], kh * kw, cin, cout).reshape(kh, kw, cin, cout).permute(3, 2, 0, 1)
        z = F.conv2d(x, wk, arr(st.b_kk[0], cout), padding=(kh // 2, kw // 2))
    z = F.gelu(z)
    if st.w_res:
        z = z + torch.einsum("bchw,cn->bnhw", x, arr(st.w_res, cin, cout)) + arr(st.b_res, cout).view(1, -1, 1, 1)
    else:
        z = z + x
    assert torch.allclose(z, ref, rtol=2e-5, atol=2e-5), (z - ref).abs().max()


def test_product_path_refuses_cpu_tensors():
    from timesnet_forecast.losses import negative_binomial_nll
    from timesnet_forecast.models.timesnet import FFTPeriodSelector, TimesBlock
    blk = TimesBlock(8, [(3, 3)], 0.0, "gelu")
    object.__setattr__(blk, "period_selector", FFTPeriodSelector(2, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        blk(torch.randn(2, 16, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        FFTPeriodSelector(2, 8)(torch.randn(2, 16, 8))
    with pytest.raises((RuntimeError, TypeError)):
        negative_binomial_nll(torch.ones(2, 3, 4), torch.ones(2, 3, 4), torch.ones(2, 3, 4))
    with pytest.raises(ValueError):
        blk(torch.randn(16, 8))


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "flow-timesnet_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "flowtimes_oracle" not in src, f"{f} references the oracle"
        # docstrings cite /root/reference/...:line; nothing may put it on sys.path or open files there
        assert not re.search(r"(sys\.path|open\(|Path\().*/root/reference", src), f"{f} reads the reference checkout"


# --------------------------------------------------------------------------- #
# N > 1: two gloo processes, batch sharded, spectrum sum all-reduced
# --------------------------------------------------------------------------- #
def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
    sys.path.insert(0, str(ROOT / "oracle"))
    import torch.distributed as dist
    import flowtimes_oracle as orc
    import flowtimes_synth as syn
    from timesnet_forecast.parallel import reduce_spectrum_sum, shard_batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, L, C, k = 8, 48, 16, 3
        x = syn.white_features(B, L, C, seed=5)              # same full batch on both ranks
        lo, hi = shard_batch(B, rank, world)
        med_local = orc.channel_median_spectrum(x[lo:hi])    # [B/world, F]
        msg = torch.cat([med_local.sum(dim=0), torch.tensor([float(hi - lo)])])   # the layout ftn_spectrum writes
        reduce_spectrum_sum(msg, None)
        total, global_b = msg[:-1], int(msg[-1].item())
        assert global_b == B
        sel = orc.select_from_spectrum(med_local, total, global_b, L, k, L - 1, 1, x.dtype)
        torch.save({"periods": sel.periods, "freq": sel.freq_indices, "mean": total / global_b, "lo": lo, "hi": hi},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_period_search_selects_identically_gloo(tmp_path):
    import torch.multiprocessing as mp
    import flowtimes_oracle as orc
    import flowtimes_synth as syn
    world = 2
    port = _free_port()
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    x = syn.white_features(8, 48, 16, seed=5)
    full = orc.select_periods(x, 3, 47, 1)
    assert (outs[0]["lo"], outs[0]["hi"], outs[1]["lo"], outs[1]["hi"]) == (0, 4, 4, 8)
    for o in outs:
        assert o["periods"].tolist() == full.periods.tolist()
        assert o["freq"].tolist() == full.freq_indices.tolist()
        assert torch.allclose(o["mean"], full.amp_mean, rtol=1e-6, atol=1e-6)
    assert torch.equal(outs[0]["mean"], outs[1]["mean"])     # bit-identical on every rank


def test_shard_batch_covers_everything():
    from timesnet_forecast.parallel import shard_batch
    for B in (1, 7, 64, 30000):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_phase_stage_images_are_toeplitz_windows():
    """w_kk_phase (tc_conv4.cu): rows [32 s, 32 s + 128) of a plane are the operand of position shift s, i.e. row
    (3 - phi) * 32 + n holds W[n, :, dr, s - phi] when 0 <= s - phi < kw and zeros otherwise."""
    import torch
    from timesnet_forecast._pack import _phase_stage_images
    g = torch.Generator().manual_seed(0)
    for kh, kw in ((3, 3), (5, 5), (7, 7)):
        wk = torch.randn(32, 32, kh, kw, generator=g, dtype=torch.float64)
        img = _phase_stage_images(wk)
        plane = (kw + 3) * 32
        assert img.shape == (kh, 4 * plane + 96, 8)
        assert img.numel() * 2 == kh * (2048 * (kw + 3) + 1536)          # bf16 bytes == c4_stage_bytes(kw) per tap row
        for dr in range(kh):
            for c in range(4):
                for s in range(kw + 3):
                    win = img[dr, c * plane + 32 * s: c * plane + 32 * s + 128]
                    for phi in range(4):
                        blk = win[(3 - phi) * 32:(4 - phi) * 32]
                        dx = s - phi
                        want = wk[:, c * 8:(c + 1) * 8, dr, dx] if 0 <= dx < kw else torch.zeros(32, 8, dtype=torch.float64)
                        assert torch.equal(blk, want), (kh, dr, c, s, phi)


def test_split_h2_is_a_22_bit_pair_with_an_exact_scale():
    """_pack.split_h2: hi + lo reproduces w * 2^s to 2^-22 relative of the largest entry, the scale is an exact power of two
    that puts the largest magnitude in [2^13, 2^14), zeros and tiny / huge tensors survive, non-finite tensors are refused."""
    import math
    import torch
    from timesnet_forecast._pack import split_h2
    g = torch.Generator().manual_seed(0)
    for mag in (1e-6, 3e-3, 0.05, 1.0, 700.0, 1e6):
        w = torch.randn(64, 48, generator=g) * mag
        w[0, 0] = 0.0
        (hi, lo), inv = split_h2(w)
        assert hi.dtype == torch.float16 and lo.dtype == torch.float16
        s = -math.log2(inv)
        assert s == int(s)                                                   # exact power of two
        amax = float(w.abs().max()) / inv
        assert 2 ** 13 <= amax < 2 ** 14
        assert bool(torch.isfinite(hi).all()) and bool(torch.isfinite(lo).all())
        back = (hi.double() + lo.double()) * inv
        err = float((back - w.double()).abs().max()) / float(w.abs().max())
        assert err < 2.0 ** -22, (mag, err)
    (hi, lo), inv = split_h2(torch.zeros(4, 4))
    assert inv == 1.0 and float(hi.abs().max()) == 0.0
    bad = torch.ones(4, 4)
    bad[1, 1] = float("inf")
    assert split_h2(bad) == (None, 0.0)


def test_row_images_put_the_taps_of_a_row_side_by_side():
    """_pack._row_images (tc_convs.cu, row mode): [kh][chunk][plane][kw][out][8 in] -- for a tap row and an 8-channel
    chunk, rows (plane, dw, n) are the N axis of ONE operand, so the kw taps of the row multiply in one MMA."""
    import torch
    from timesnet_forecast._pack import _row_images, _tap_images, _tap_images_h2
    g = torch.Generator().manual_seed(1)
    for kh, kw in ((3, 3), (3, 5), (7, 7), (1, 7)):
        wk = torch.randn(16, 16, kh, kw, generator=g, dtype=torch.float64)
        img3 = _tap_images(wk)                                               # [tap][chunk][3][out][8]
        row = _row_images(img3[:, :, :1], kh, kw)
        assert tuple(row.shape) == (kh, 2, 1, kw, 16, 8)
        for dr in range(kh):
            for dw in range(kw):
                assert torch.equal(row[dr, :, 0, dw], img3[dr * kw + dw, :, 0])
                want = wk[:, :, dr, dw].to(torch.float32).to(torch.bfloat16)       # [out][in]
                got = row[dr, :, 0, dw].permute(1, 0, 2).reshape(16, 16)             # [out][chunk * 8 + in]
                assert torch.equal(got, want)
        img2, inv = _tap_images_h2(wk)
        row2 = _row_images(img2, kh, kw)
        assert tuple(row2.shape) == (kh, 2, 2, kw, 16, 8) and row2.dtype == torch.float16
        back = (row2[:, :, 0].double() + row2[:, :, 1].double()) * inv       # [kh][chunk][kw][out][8]
        want = wk.permute(2, 3, 0, 1).reshape(kh, kw, 16, 2, 8).permute(0, 3, 1, 2, 4)   # [kh][chunk][kw][out][8]
        assert float((back - want).abs().max()) < 2.0 ** -20 * float(wk.abs().max())
