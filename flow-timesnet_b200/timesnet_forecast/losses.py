"""Negative-Binomial likelihood (drop-in for timesnet_forecast/losses.py).

``negative_binomial_nll`` runs as one fused elementwise + reduction pass in
libflowtimes (lgamma / log1p in fp32 device math, deterministic two-level
reduction).  When ``rate`` or ``dispersion`` require grad the call goes through
``timesnet_forecast.autograd.nb_nll`` and is differentiable w.r.t. both
(``ftn_nb_nll_backward``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as nv


def _expand_mask(mask: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    mb = mask.to(dtype=torch.bool)
    if mb.ndim < like.ndim:
        mb = mb.reshape(*mb.shape, *([1] * (like.ndim - mb.ndim)))
    return mb.expand_as(like)


def negative_binomial_mask(y: torch.Tensor, rate: torch.Tensor, dispersion: torch.Tensor,
                           mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Boolean validity mask (losses.py:6-24): finite everywhere and user mask set.
    Pure bookkeeping on the caller's device; the NLL kernel recomputes it in-register."""
    valid = torch.isfinite(y) & torch.isfinite(rate) & torch.isfinite(dispersion)
    if mask is not None:
        valid = valid & _expand_mask(mask, valid)
    return valid


def negative_binomial_nll(y: torch.Tensor, rate: torch.Tensor, dispersion: torch.Tensor,
                          mask: Optional[torch.Tensor] = None, eps: float = 1e-8) -> torch.Tensor:
    """Masked mean NB negative log-likelihood, fp32 scalar (losses.py:27-58)."""
    nv.require_cuda(rate, "rate")
    if not (y.shape == rate.shape == dispersion.shape):
        y, rate, dispersion = torch.broadcast_tensors(y, rate, dispersion)
    needs_grad = torch.is_grad_enabled() and (rate.requires_grad or dispersion.requires_grad)
    with torch.no_grad():
        yf = nv.require_cuda(y.to(device=rate.device), "y").to(torch.float32).contiguous()
        m8 = None
        if mask is not None:
            m8 = _expand_mask(mask.to(rate.device), yf).to(torch.uint8).contiguous()
        if not needs_grad:
            rf = rate.detach().to(torch.float32).contiguous()
            df = nv.require_cuda(dispersion, "dispersion").detach().to(torch.float32).contiguous()
            return nv.nb_nll(yf, rf, df, m8, eps)
    from .autograd import nb_nll as _nb_nll_grad
    return _nb_nll_grad(yf, rate, nv.require_cuda(dispersion, "dispersion"), m8, eps)
