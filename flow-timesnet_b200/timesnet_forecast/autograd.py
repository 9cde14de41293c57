"""Backward pass, first slice: differentiable wrappers around the stages where a backward is
cheapest -- the NB negative log-likelihood, the NB head (time projection + mu / sigma heads + softplus) and LayerNorm.

Every op, forward and backward, runs in libflowtimes kernels (``ftn_nb_nll_backward``,
``ftn_nb_head_epilogue_backward``, ``ftn_layer_norm_backward``, ``ftn_gemm_f32``); torch only provides
``autograd.Function`` bookkeeping.  The Inception chain and the period selector have no backward yet, so
``TimesBlock`` / ``TimesNet`` stay forward-only (they raise on inputs that require grad); these functions are the
building blocks the training path will be assembled from, and are checked against float64 autograd of the reference
formulas in tests/test_gpu_backward.py.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as nv


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return nv.require_cuda(t, "tensor").detach().to(torch.float32).contiguous()


def _ones(n: int, device) -> torch.Tensor:
    return torch.ones(1, n, dtype=torch.float32, device=device)


def _colsum(m2d: torch.Tensor) -> torch.Tensor:
    """Column sums of ``[rows, cols]`` as a GEMM with a row of ones (deterministic, native)."""
    return nv.gemm_f32(_ones(m2d.shape[0], m2d.device), m2d).reshape(-1)


class _NBNLL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, rate, disp, mask_u8, eps):
        yf, rf, df = _f32c(y), _f32c(rate), _f32c(disp)
        ctx.save_for_backward(yf, rf, df, mask_u8 if mask_u8 is not None else torch.empty(0))
        ctx.eps = float(eps)
        ctx.has_mask = mask_u8 is not None
        return nv.nb_nll(yf, rf, df, mask_u8, eps)

    @staticmethod
    def backward(ctx, grad_out):
        yf, rf, df, m8 = ctx.saved_tensors
        d_rate, d_disp = nv.nb_nll_backward(yf, rf, df, m8 if ctx.has_mask else None, ctx.eps, grad_out)
        return None, d_rate, d_disp, None, None


def nb_nll(y: torch.Tensor, rate: torch.Tensor, dispersion: torch.Tensor, mask_u8: Optional[torch.Tensor] = None,
           eps: float = 1e-8) -> torch.Tensor:
    """Differentiable NB-NLL (losses.py:27-58): gradients w.r.t. ``rate`` and ``dispersion``."""
    return _NBNLL.apply(y, rate, dispersion, mask_u8, eps)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        xf, wf, bf = _f32c(x), _f32c(weight), _f32c(bias)
        ctx.save_for_backward(xf, wf)
        ctx.eps = float(eps)
        return nv.layer_norm(xf, wf, bf, eps)

    @staticmethod
    def backward(ctx, dy):
        xf, wf = ctx.saved_tensors
        dx, dw, db = nv.layer_norm_backward(xf, _f32c(dy), wf, ctx.eps)
        return dx, dw, db, None


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """Differentiable fp32 LayerNorm over the last dim (timesnet.py:1162-1181)."""
    return _LayerNorm.apply(x, weight, bias, eps)


class _NBHead(torch.autograd.Function):
    """rate, dispersion = NB head(seq) (timesnet.py:2063-2093), gradients w.r.t. seq, forecast_time_proj, mu_head,
    sigma_head, the late-bias input and its gate."""

    @staticmethod
    def forward(ctx, seq, Wt, bt, Wmu, bmu, Wsg, bsg, hist, late, late_gate, floor_n):
        seq, Wt, bt, Wmu, bmu, Wsg, bsg, hist, floor_n = (_f32c(t) for t in (seq, Wt, bt, Wmu, bmu, Wsg, bsg, hist, floor_n))
        late = None if late is None else _f32c(late)
        late_gate = None if late_gate is None else _f32c(late_gate).reshape(-1)
        steps, N = Wt.shape[0], Wmu.shape[0]
        flags = torch.zeros(1, dtype=torch.int32, device=seq.device)
        rate, disp = nv.nb_head(seq, steps, N, Wt, bt, Wmu, bmu, Wsg, bsg, hist, late, late_gate, floor_n, flags)
        ctx.save_for_backward(seq, Wt, bt, Wmu, Wsg, floor_n, rate, disp,
                              late if late is not None else torch.empty(0),
                              late_gate if late_gate is not None else torch.empty(0))
        ctx.has_late = late is not None
        return rate, disp

    @staticmethod
    def backward(ctx, d_rate, d_disp):
        seq, Wt, bt, Wmu, Wsg, floor_n, rate, disp, late, gate = ctx.saved_tensors
        B, L, C = seq.shape
        steps, N = Wt.shape[0], Wmu.shape[0]
        dpr, dpd = nv.nb_head_epilogue_backward(rate, disp, floor_n, _f32c(d_rate), _f32c(d_disp))   # [B, steps, N]
        dpr2, dpd2 = dpr.reshape(B * steps, N), dpd.reshape(B * steps, N)
        hidden = nv.gemm_f32(Wt, seq) + bt.view(1, steps, 1)                 # recomputed: [B, steps, C]
        hid2 = hidden.reshape(B * steps, C).contiguous()
        dWmu = nv.gemm_f32(dpr2, hid2, trans_a=True)                         # [N, C]
        dWsg = nv.gemm_f32(dpd2, hid2, trans_a=True)
        dbmu, dbsg = _colsum(dpr2), _colsum(dpd2)
        d_hidden = (nv.gemm_f32(dpr2, Wmu) + nv.gemm_f32(dpd2, Wsg)).reshape(B, steps, C).contiguous()
        d_seq = nv.gemm_f32(Wt, d_hidden, trans_a=True)                      # Wt^T [L, steps] . d_hidden[b] -> [B, L, C]
        dWt_b = nv.gemm_f32(d_hidden, seq, trans_b=True)                     # [B, steps, L]
        dWt = _colsum(dWt_b.reshape(B, steps * L)).reshape(steps, L)
        dbt = _colsum(d_hidden.permute(0, 2, 1).reshape(B * C, steps).contiguous())
        d_late = d_gate = None
        if ctx.has_late:                                                     # pre += gate[h] * late[b, n, h]
            d_late = (dpr * gate.view(1, steps, 1)).permute(0, 2, 1).contiguous()
            d_gate = _colsum((dpr.permute(0, 2, 1) * late).reshape(B * N, steps).contiguous())
        return d_seq, dWt, dbt, dWmu, dbmu, dWsg, dbsg, None, d_late, d_gate, None


def nb_head(seq, Wt, bt, Wmu, bmu, Wsg, bsg, hist, late=None, late_gate=None, floor_n=None):
    """Differentiable NB head.  ``seq`` fp32 ``[B, L, C]``; ``Wt`` ``[steps, L]``; heads ``[N, C]``; ``hist`` ``[B, steps, N]``;
    ``late`` ``[B, N, steps]`` with ``late_gate`` ``[steps]`` or both ``None``; ``floor_n`` ``[N]``."""
    return _NBHead.apply(seq, Wt, bt, Wmu, bmu, Wsg, bsg, hist, late, late_gate, floor_n)


# --------------------------------------------------------------------------- #
# second slice: the pieces of the Inception chain and of the aggregation
# --------------------------------------------------------------------------- #
class _Act(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act_code):
        xf = _f32c(x)
        ctx.save_for_backward(xf)
        ctx.act = int(act_code)
        return nv.act_forward(xf, ctx.act)

    @staticmethod
    def backward(ctx, dy):
        (xf,) = ctx.saved_tensors
        return nv.act_backward(xf, _f32c(dy), ctx.act), None


def activation(x: torch.Tensor, name: str = "gelu") -> torch.Tensor:
    """Exact-erf GELU (``nn.GELU()`` default, timesnet.py:643) or ReLU, forward and backward in libflowtimes."""
    return _Act.apply(x, nv.FTN_ACT_RELU if name.lower() == "relu" else nv.FTN_ACT_GELU)


class _Conv2dGrid(torch.autograd.Function):
    """``F.conv2d(x, w, b, padding="same")`` (odd kernel) on an NCHW grid, computed on the zero-copy fold
    ``[B, H*W, C]`` by ``ftn_conv2d_grid``.  Backward: the data gradient is the same kernel with the taps flipped and
    cin / cout swapped, the weight gradient ``ftn_conv2d_grid_backward_weight``, the bias gradient a column sum."""

    @staticmethod
    def forward(ctx, x_nchw, weight, bias):
        B, Cin, H, W = x_nchw.shape
        Cout, _, kh, kw = weight.shape
        xs = _f32c(x_nchw).permute(0, 2, 3, 1).reshape(B, H * W, Cin).contiguous()        # fold: memory plumbing only
        w_taps = _f32c(weight).permute(2, 3, 1, 0).reshape(kh * kw, Cin, Cout).contiguous()
        b = _f32c(bias) if bias is not None else torch.zeros(Cout, dtype=torch.float32, device=xs.device)
        plan = nv.single_group_plan(W, H, xs.device)
        out = nv.conv2d_grid(xs, plan, w_taps, b, kh, kw)
        ctx.save_for_backward(xs, w_taps)
        ctx.geom = (B, Cin, Cout, H, W, kh, kw, bias is not None)
        ctx.plan = plan
        return out.view(B, H, W, Cout).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy_nchw):
        xs, w_taps = ctx.saved_tensors
        B, Cin, Cout, H, W, kh, kw, has_bias = ctx.geom
        dy = _f32c(dy_nchw).permute(0, 2, 3, 1).reshape(B, H * W, Cout).contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            w_back = w_taps.flip(0).transpose(1, 2).contiguous()                         # tap (dr, dw) -> (kh-1-dr, kw-1-dw)
            zero = torch.zeros(Cin, dtype=torch.float32, device=dy.device)
            dxs = nv.conv2d_grid(dy, ctx.plan, w_back, zero, kh, kw)
            dx = dxs.view(B, H, W, Cin).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dwt = nv.conv2d_grid_backward_weight(xs, dy, W, kh, kw)                       # [taps, cin, cout]
            dw = dwt.view(kh, kw, Cin, Cout).permute(3, 2, 0, 1)
        if has_bias and ctx.needs_input_grad[2]:
            db = _colsum(dy.view(B * H * W, Cout))
        return dx, dw, db


def conv2d_same(x_nchw: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Differentiable ``conv2d`` with zero "same" padding and an odd kernel on an NCHW grid (InceptionBranch's convs,
    timesnet.py:575-593)."""
    return _Conv2dGrid.apply(x_nchw, weight, bias)


def inception_branch(x_nchw: torch.Tensor, branch: torch.nn.Sequential) -> torch.Tensor:
    """Differentiable ``InceptionBranch.forward`` (timesnet.py:592-593): the branch's convs in order."""
    out = x_nchw
    for conv in branch:
        out = conv2d_same(out, conv.weight, conv.bias)
    return out


def inception_block(x_nchw: torch.Tensor, block) -> torch.Tensor:
    """Differentiable ``InceptionBlock.forward`` (timesnet.py:645-654): ``act(proj(cat_j branch_j(x))) + res_proj(x)``
    (dropout is the identity: the build has no RNG on the path; ``dropout > 0`` in train mode raises upstream)."""
    feats = [inception_branch(x_nchw, p.branch) for p in block.paths]
    cat = torch.cat(feats, dim=1)                                                        # memory plumbing
    out = conv2d_same(cat, block.proj.weight, block.proj.bias)
    out = activation(out, "relu" if isinstance(block.act, torch.nn.ReLU) else "gelu")
    res = x_nchw if isinstance(block.res_proj, torch.nn.Identity) else conv2d_same(x_nchw, block.res_proj.weight,
                                                                                    block.res_proj.bias)
    return out + res


class _Aggregate(torch.autograd.Function):
    """``x + sum_g w[b, g] * delta_g`` (timesnet.py:1075-1099, :818) for fp32 tensors and a device plan."""

    @staticmethod
    def forward(ctx, x, delta, weights, plan_dev):
        xf, df, wf = _f32c(x), _f32c(delta), _f32c(weights)
        out = torch.empty_like(xf)
        nv.aggregate(xf, df, wf, plan_dev, None, None, 0.0, out)
        ctx.save_for_backward(df, wf, plan_dev)
        return out

    @staticmethod
    def backward(ctx, d_out):
        df, wf, plan_dev = ctx.saved_tensors
        g = _f32c(d_out)
        d_delta, d_w = nv.aggregate_backward(g, df, wf, plan_dev)
        return g, d_delta, d_w, None


def aggregate(x: torch.Tensor, delta: torch.Tensor, weights: torch.Tensor, plan_dev: torch.Tensor) -> torch.Tensor:
    """Differentiable softmax-weighted aggregation + residual given the group weights ``[B, 16]`` and the device plan."""
    return _Aggregate.apply(x, delta, weights, plan_dev)


class _GroupWeights(torch.autograd.Function):
    """Group weights ``[B, 16]`` of a plan from per-window amplitudes ``[B, k]`` (softmax over the valid candidates,
    scatter-added per group; timesnet.py:992-1009).  The forward value is the one the search already produced."""

    @staticmethod
    def forward(ctx, amps, weights_value, plan_dev):
        ctx.save_for_backward(_f32c(amps), plan_dev)
        return weights_value.detach().clone()

    @staticmethod
    def backward(ctx, d_w):
        amps, plan_dev = ctx.saved_tensors
        return nv.group_weights_backward(amps, plan_dev, _f32c(d_w)), None, None


class _SpectrumAmps(torch.autograd.Function):
    """Per-window amplitudes at the selected bins, ``median_c |rfft_t x|[f_j]`` (timesnet.py:109-111, :134), as a function
    of x: the value is the search's, the backward sends each amplitude's gradient to its lower-median channel."""

    @staticmethod
    def forward(ctx, x, amps_value, plan_dev):
        ctx.save_for_backward(_f32c(x), plan_dev)
        return amps_value.detach().float().clone()

    @staticmethod
    def backward(ctx, d_amps):
        xf, plan_dev = ctx.saved_tensors
        return nv.spectrum_amp_backward(xf, plan_dev, _f32c(d_amps)), None, None


def period_weights(x: torch.Tensor, plan) -> torch.Tensor:
    """The plan's group weights as a differentiable function of x (FFT selector) -- the second path from the
    aggregation back into the input that the reference's autograd follows."""
    amps = _SpectrumAmps.apply(x, plan.amps, plan.plan_dev)
    return _GroupWeights.apply(amps, plan.weights, plan.plan_dev)


def group_weights(amps: torch.Tensor, plan) -> torch.Tensor:
    """Same for amplitudes a custom selector module returned (they may carry their own graph)."""
    return _GroupWeights.apply(amps, plan.weights, plan.plan_dev)


class _Linear(torch.autograd.Function):
    """``F.linear`` on the last dim (fp32): forward ``ftn_linear``, backward the two GEMMs dX = dY . W, dW = dY^T . X
    (``ftn_gemm_f32``) and a column sum for the bias."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xf, wf = _f32c(x), _f32c(weight)
        bf = None if bias is None else _f32c(bias)
        ctx.save_for_backward(xf, wf)
        ctx.has_bias = bias is not None
        return nv.linear(xf, wf, bf)

    @staticmethod
    def backward(ctx, dy):
        xf, wf = ctx.saved_tensors
        K, N = xf.shape[-1], wf.shape[0]
        dy2 = _f32c(dy).reshape(-1, N)
        x2 = xf.reshape(-1, K)
        dx = nv.gemm_f32(dy2, wf).reshape(xf.shape) if ctx.needs_input_grad[0] else None
        dw = nv.gemm_f32(dy2, x2, trans_a=True) if ctx.needs_input_grad[1] else None
        db = _colsum(dy2) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable fp32 ``nn.Linear`` (value / temporal embeddings, static and context projections, late-bias head)."""
    return _Linear.apply(x, weight, bias)
