"""One-time host-side packing of InceptionBlock parameters for libflowtimes.

The reference keeps 3 x (1x1 -> k x k -> 1x1) convs + a 1x1 ``proj`` over their
concatenation (timesnet.py:586-590, :633, :650-651).  There is no activation
between a branch's last 1x1 and ``proj``, so they compose exactly:

    proj(cat_j(W3_j h_j + b3_j)) = sum_j (P_j W3_j) h_j + (b_p + sum_j P_j b3_j)

with ``P_j`` the j-th column block of ``proj.weight``.  The packed form therefore
is: one concatenated input 1x1 (cin -> n_branch*mid), the per-branch k x k convs,
and ONE folded output 1x1 (n_branch*mid -> cout).  For bottleneck_ratio == 1
(single k x k conv per branch, timesnet.py:575-580) ``proj`` folds into a single
conv with the union (largest) window.  Folding is done in float64 and costs
~1e-6 relative.  Layouts are K-major as
include/flowtimes.h describes.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Tuple

import torch
from torch import nn

from . import _native as nv


class PackedInception:
    """Owns the packed device tensors and the C struct that points at them."""

    def __init__(self, struct: nv.FtnInceptionWeights, tensors: List[torch.Tensor], executed_macs: int):
        self.struct = struct
        self.tensors = tensors            # keep-alive
        self.executed_macs_per_pos = executed_macs


def _conv_list(branch_seq: nn.Sequential) -> List[nn.Conv2d]:
    convs = [m for m in branch_seq if isinstance(m, nn.Conv2d)]
    if len(convs) not in (1, 3):
        raise ValueError("InceptionBranch must hold 1 (ratio 1) or 3 (bottleneck) Conv2d modules")
    return convs


def pack_inception_block(block: nn.Module, device: torch.device) -> PackedInception:
    paths = [_conv_list(p.branch) for p in block.paths]
    n_branch = len(paths)
    if n_branch > nv.FTN_MAX_BRANCH:
        raise ValueError(f"kernel_set has {n_branch} entries; libflowtimes supports at most {nv.FTN_MAX_BRANCH}")
    depth = {len(p) for p in paths}
    if len(depth) != 1:
        raise ValueError("all inception branches must share the same bottleneck structure")
    bottleneck = depth.pop() == 3
    f64 = torch.float64
    proj_w = block.proj.weight.detach().to("cpu", f64)[:, :, 0, 0]          # [cout, n_branch*cout]
    proj_b = block.proj.bias.detach().to("cpu", f64)
    cout = proj_w.shape[0]
    cin = paths[0][0].weight.shape[1]
    st = nv.FtnInceptionWeights()
    keep: List[torch.Tensor] = []

    def dev(t: torch.Tensor) -> int:
        d = t.to(torch.float32).contiguous().to(device)
        keep.append(d)
        return d.data_ptr()

    def dev16(t: torch.Tensor) -> int:
        """bf16 copy in [N][K] (K contiguous) layout for the tcgen05 path."""
        d = t.to(torch.float32).to(torch.bfloat16).contiguous().to(device)
        keep.append(d)
        return d.data_ptr()

    def dev_planes(planes: torch.Tensor) -> int:
        """already-bf16 tensor of split planes -> device"""
        d = planes.contiguous().to(device)
        keep.append(d)
        return d.data_ptr()

    def dev_h2(t: torch.Tensor) -> Tuple[int, float]:
        """[N][K] weight -> two fp16 planes [N][2 K] of (t * 2^s) on the device, and 2^-s (split_h2 below)"""
        planes, inv = split_h2(t)
        if planes is None:
            return None, 0.0
        d = torch.cat(planes, dim=1).contiguous().to(device)
        keep.append(d)
        return d.data_ptr(), inv

    st.cin, st.cout, st.n_branch = int(cin), int(cout), n_branch
    macs = 0
    if bottleneck:
        mid = paths[0][0].weight.shape[0]
        st.mid = int(mid)
        st.kk_cin = st.kk_cout = int(mid)
        w_in = torch.cat([p[0].weight.detach().to("cpu", f64)[:, :, 0, 0] for p in paths], dim=0)   # [NB, cin]
        b_in = torch.cat([p[0].bias.detach().to("cpu", f64) for p in paths], dim=0)
        st.w_in, st.b_in = dev(w_in.t()), dev(b_in)
        st.w_in_bf16 = dev16(w_in)                                           # [NB][cin]
        tc_ok = mid % 16 == 0 and cin % 16 == 0 and cout % 16 == 0           # tensor-core tile granularity
        if tc_ok:
            st.w_in_s3 = dev_planes(torch.cat(split3(w_in), dim=1))          # [NB][3 cin]
            st.w_in_h2, st.sc_in = dev_h2(w_in)                              # [NB][2 cin] fp16
        macs += cin * n_branch * mid
        w_out_rows = []
        b_out = proj_b.clone()
        for j, p in enumerate(paths):
            wk = p[1].weight.detach().to("cpu", f64)                         # [mid_out, mid_in, kh, kw]
            kh, kw = int(wk.shape[2]), int(wk.shape[3])
            st.kh[j], st.kw[j] = kh, kw
            st.w_kk[j] = dev(wk.permute(2, 3, 1, 0).reshape(kh * kw, mid, mid))
            st.w_kk_bf16[j] = dev16(wk.permute(2, 3, 0, 1).reshape(kh * kw, mid, mid))   # [tap][out][in]
            st.b_kk[j] = dev(p[1].bias.detach().to("cpu", f64))
            if mid == 32 and kh % 2 == 1 and kw % 2 == 1:
                st.w_kk_phase[j] = dev16(_phase_stage_images(wk))
            if tc_ok and kh % 2 == 1 and kw % 2 == 1:
                img3 = _tap_images(wk)                                       # [tap][mid/8][3][mid][8] bf16
                st.w_kk_img3[j] = dev_planes(img3)
                st.w_kk_img[j] = dev_planes(img3[:, :, 0])
                img2, inv = _tap_images_h2(wk)                               # [tap][mid/8][2][mid][8] fp16
                if img2 is not None:
                    st.w_kk_img2[j] = dev_planes(img2)
                    st.sc_kk[j] = inv
                if kw * mid <= 256:                                          # row mode of tc_convs (narrow branches)
                    st.w_kk_row[j] = dev_planes(_row_images(img3[:, :, :1], kh, kw))
                    if img2 is not None and 2 * kw * mid <= 256:
                        st.w_kk_row2[j] = dev_planes(_row_images(img2, kh, kw))
            macs += kh * kw * mid * mid
            P = proj_w[:, j * cout:(j + 1) * cout]                           # [cout, cout]
            W3 = p[2].weight.detach().to("cpu", f64)[:, :, 0, 0]             # [cout, mid]
            w_out_rows.append((P @ W3).t())                                  # [mid, cout]
            b_out = b_out + P @ p[2].bias.detach().to("cpu", f64)
        w_out_kn = torch.cat(w_out_rows, dim=0)                              # [NB][cout]
        st.w_out, st.b_out = dev(w_out_kn), dev(b_out)
        st.w_out_bf16 = dev16(w_out_kn.t())                                  # [cout][NB]
        if tc_ok:
            st.w_out_s3 = dev_planes(torch.cat(split3(w_out_kn.t()), dim=1))  # [cout][3 NB]
            st.w_out_h2, st.sc_out = dev_h2(w_out_kn.t())                     # [cout][2 NB] fp16
        macs += n_branch * mid * cout
    else:
        st.mid = 0
        st.kk_cin, st.kk_cout = int(cin), int(cout)
        KH = max(int(p[0].weight.shape[2]) for p in paths)
        KW = max(int(p[0].weight.shape[3]) for p in paths)
        wtot = torch.zeros(cout, cin, KH, KW, dtype=f64)
        btot = proj_b.clone()
        for j, p in enumerate(paths):
            w = p[0].weight.detach().to("cpu", f64)                          # [cout, cin, kh, kw]
            kh, kw = int(w.shape[2]), int(w.shape[3])
            if (KH - kh) % 2 or (KW - kw) % 2:
                raise ValueError("kernel_set sizes must share parity to share a centred window")
            P = proj_w[:, j * cout:(j + 1) * cout]
            r0, s0 = (KH - kh) // 2, (KW - kw) // 2
            wtot[:, :, r0:r0 + kh, s0:s0 + kw] += torch.einsum("om,mirs->oirs", P, w)
            btot = btot + P @ p[0].bias.detach().to("cpu", f64)
        st.n_branch = 1
        st.kh[0], st.kw[0] = KH, KW
        st.w_kk[0] = dev(wtot.permute(2, 3, 1, 0).reshape(KH * KW, cin, cout))
        st.b_kk[0] = dev(btot)
        macs += KH * KW * cin * cout
    if isinstance(block.res_proj, nn.Conv2d):
        st.w_res = dev(block.res_proj.weight.detach().to("cpu", f64)[:, :, 0, 0].t())   # [cin, cout]
        st.w_res_bf16 = dev16(block.res_proj.weight.detach().to("cpu", f64)[:, :, 0, 0])  # [cout][cin]
        st.b_res = dev(block.res_proj.bias.detach().to("cpu", f64))
        if bottleneck and tc_ok:
            st.w_res_s3 = dev_planes(torch.cat(split3(block.res_proj.weight.detach().to("cpu", f64)[:, :, 0, 0]), dim=1))
            st.w_res_h2, st.sc_res = dev_h2(block.res_proj.weight.detach().to("cpu", f64)[:, :, 0, 0])
        macs += cin * cout
    else:
        st.w_res, st.b_res = None, None
    st.w_mid_first, st.w_mid_second = None, None
    if bottleneck and isinstance(block.res_proj, nn.Conv2d):
        # stage images of the fused middle kernel (tc_mid.cu), see include/flowtimes.h
        NB = n_branch * int(st.mid)
        w_res_nk = block.res_proj.weight.detach().to("cpu", f64)[:, :, 0, 0]          # [cout][cin]
        w_out_nk = w_out_kn.t()                                                        # [cout][NB]

        def kblocks(w_nk: torch.Tensor) -> torch.Tensor:                               # [N][K] -> [N/128][kb][128][64]
            N_, K_ = w_nk.shape
            kb = (K_ + 63) // 64
            padded = torch.zeros(N_, kb * 64, dtype=f64)
            padded[:, :K_] = w_nk
            return padded.reshape(N_ // 128, 128, kb, 64).permute(0, 2, 1, 3)

        if cout % 128 == 0:
            first = torch.cat([kblocks(w_out_nk), kblocks(w_res_nk)], dim=1)           # [cout/128][kb1+kb2][128][64]
            st.w_mid_first = dev16(first)
        if cin % 128 == 0:
            # halved: tc_mid feeds this stage 2 * a2 (its second activation is evaluated as 2 * act, one instruction
            # less); a power of two, so the products are bit-identical to a2 . w
            both = 0.5 * torch.cat([w_in, w_res_nk], dim=0)                            # [NB+cout][cin]
            second = both.reshape(NB + cout, cin // 64, 64).permute(1, 0, 2)           # [cin/64][NB+cout][64]
            st.w_mid_second = dev16(second)                                            # == [cin/128][2][NB+cout][64]
    return PackedInception(st, keep, macs)


def split3(t: torch.Tensor):
    """fp32 value of ``t`` as three bf16 planes (hi, mid, lo): hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid).
    ``t`` is first rounded to fp32 (the reference's parameter dtype); the residuals are exact in fp32."""
    v = t.to(torch.float32)
    hi = v.to(torch.bfloat16)
    r1 = v - hi.to(torch.float32)
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.to(torch.float32)).to(torch.bfloat16)
    return hi, mid, lo


def split_h2(t: torch.Tensor):
    """fp32 value of ``t`` as TWO fp16 planes of ``t * 2^s``: hi = fp16(v), lo = fp16(v - hi), 22 significand bits.
    ``s`` brings the largest magnitude into [2^13, 2^14) so that the lo plane of typical weights stays out of the fp16
    subnormals; the kernels multiply their accumulators by the returned ``2^-s`` (exact).  Returns ``(None, 0.0)`` when
    the tensor has non-finite entries (the three-plane bf16 form then runs, it has fp32's range)."""
    v = t.to(torch.float32)
    if not bool(torch.isfinite(v).all()):
        return None, 0.0
    amax = float(v.abs().max())
    s = 0 if amax == 0.0 else int(math.floor(math.log2(16383.0 / amax)))
    s = max(-100, min(100, s))
    scaled = v * (2.0 ** s)                                                  # exact: a power of two
    hi = scaled.to(torch.float16)
    lo = (scaled - hi.to(torch.float32)).to(torch.float16)
    return (hi, lo), 2.0 ** (-s)


def _tap_images_h2(wk: torch.Tensor):
    """fp16 two-plane form of ``_tap_images``: [kh*kw][mid/8 chunks][2 planes][mid out][8 in] and the inverse scale."""
    n_out, n_in, kh, kw = (int(v) for v in wk.shape)
    planes, inv = split_h2(wk)
    if planes is None:
        return None, 0.0
    st = torch.stack(planes, dim=0)                                          # [2][out][in][kh][kw]
    img = st.permute(3, 4, 0, 2, 1).reshape(kh * kw, 2, n_in // 8, 8, n_out)  # [tap][plane][chunk][8 in][out]
    return img.permute(0, 2, 1, 4, 3).contiguous(), inv                       # [tap][chunk][plane][out][8 in]


def _row_images(img: torch.Tensor, kh: int, kw: int) -> torch.Tensor:
    """Per-tap images ``[kh*kw][chunk][plane][out][8]`` -> per-tap-ROW images ``[kh][chunk][plane][kw][out][8]``: the kw
    taps of a row side by side on N (tc_convs.cu, row mode)."""
    taps, chunks, planes, n_out, eight = (int(v) for v in img.shape)
    assert taps == kh * kw
    r = img.reshape(kh, kw, chunks, planes, n_out, eight).permute(0, 2, 3, 1, 4, 5)
    return r.contiguous()


def _tap_images(wk: torch.Tensor) -> torch.Tensor:
    """Per-tap weight images of the streaming k x k kernel (tc_convs.cu; layout in include/flowtimes.h).

    wk: [mid out, mid in, kh, kw] -> bf16 [kh*kw][mid/8 chunks][3 planes][mid out][8 in]: per 8-channel chunk the
    three weight planes are consecutive row blocks, i.e. ONE K-major operand with N = 3 mid rows."""
    n_out, n_in, kh, kw = (int(v) for v in wk.shape)
    planes = torch.stack(split3(wk), dim=0)                                  # [3][out][in][kh][kw]
    img = planes.permute(3, 4, 0, 2, 1).reshape(kh * kw, 3, n_in // 8, 8, n_out)   # [tap][plane][chunk][8 in][out]
    return img.permute(0, 2, 1, 4, 3).contiguous()                           # [tap][chunk][plane][out][8 in]


def _phase_stage_images(wk: torch.Tensor) -> torch.Tensor:
    """Stage images of the phases-on-M k x k kernel (tc_conv4.cu; layout in include/flowtimes.h).

    wk: [32 out, 32 in, kh, kw].  Per tap row: four 8-input-channel planes of (kw + 3) * 32 rows, each
    starting with 3 zero blocks of 32 rows, plus 3 closing zero blocks -- any 128-row window of a plane
    is the Toeplitz operand of one (tap row, position shift) pair.
    """
    n_out, n_in, kh, kw = (int(v) for v in wk.shape)
    plane = (kw + 3) * 32
    img = torch.zeros(kh, (n_in // 8) * plane + 96, 8, dtype=wk.dtype)
    blk = wk.permute(2, 1, 3, 0).reshape(kh, n_in // 8, 8, kw, n_out).permute(0, 1, 3, 4, 2)   # [kh][c][dx][n][8]
    for c in range(n_in // 8):
        img[:, c * plane + 96:c * plane + 96 + kw * 32] = blk[:, c].reshape(kh, kw * n_out, 8)
    return img


def params_fingerprint(module: nn.Module) -> Tuple:
    return tuple((p.data_ptr(), p._version, p.device.type) for p in module.parameters())


def split_linear_weight(w: torch.Tensor, k_pad: int, rows_pad: int = 0, row_offset: int = 0) -> torch.Tensor:
    """``nn.Linear`` weight ``[N, K]`` as the three-plane bf16 operand of the tensor-core row GEMMs: ``[rows][3 k_pad]``
    (plane p at columns ``[p k_pad, (p + 1) k_pad)``, K zero-padded to ``k_pad``); with ``rows_pad`` the N rows are
    placed at ``row_offset`` of a zero matrix with ``rows_pad`` rows."""
    N, K = w.shape
    rows = rows_pad if rows_pad > 0 else N
    out = torch.zeros(rows, 3, k_pad, dtype=torch.bfloat16, device=w.device)
    for p, plane in enumerate(split3(w.detach())):
        out[row_offset:row_offset + N, p, :K] = plane
    return out.reshape(rows, 3 * k_pad).contiguous()
