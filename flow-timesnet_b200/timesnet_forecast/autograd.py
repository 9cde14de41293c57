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
