"""Forecast helpers of the reference's predict.py that sit on the hot path
(predict.py:261-342).  The CSV / artifact pipeline around them is out of scope.

``forecast_recursive_batch`` keeps the reference signature.  For the B200
``TimesNet`` the rolling loop is DEVICE RESIDENT: the window, the time marks, the
``[B, H, N]`` outputs and the step index live on the GPU and one kernel
(``ftn_recursive_advance``) does ``append -> roll`` in place after each forward, so
no step needs the host.  ``RecursiveForecaster`` additionally captures one step
(forward + advance) in a CUDA graph and replays it ``H`` times -- the 28 rolling
steps of BASELINE config 5 become 28 graph launches without a host sync.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _native as nv


def _invoke_model(model, xb, *, x_mark=None, series_static=None, series_ids=None):
    kwargs = {}
    if x_mark is not None:
        kwargs["x_mark"] = x_mark
    if series_static is not None:
        kwargs["series_static"] = series_static
    if series_ids is not None:
        kwargs["series_ids"] = series_ids
    return model(xb, **kwargs)


def forecast_direct_batch(model, last_seq: torch.Tensor, x_mark: Optional[torch.Tensor] = None,
                          series_static: Optional[torch.Tensor] = None,
                          series_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    return _invoke_model(model, last_seq, x_mark=x_mark, series_static=series_static, series_ids=series_ids)


def _is_native_model(model) -> bool:
    from .models.timesnet import TimesNet
    return isinstance(model, TimesNet)


def _check_marks(x_mark, y_mark, H: int) -> None:
    if x_mark is not None:
        if y_mark is None:
            raise ValueError(
                "Temporal features provided for history but missing future marks during recursive forecast")
        if y_mark.size(1) < H:
            raise ValueError("y_mark does not provide enough future steps for recursive forecasting")


class RecursiveForecaster:
    """Device-resident rolling one-step forecast of a B200 ``TimesNet`` (predict.py:307-342).

    State on the device: ``window[B, L, N]`` (the last ``input_len`` steps -- all a forward reads,
    timesnet.py:1877), optional ``mark[B, L, Tm]`` / ``y_mark[B, H, Tm]``, outputs ``rates / disps[B, H, N]`` and an
    int32 step counter.  ``graph=True`` captures (forward + ``ftn_recursive_advance``) once and replays it ``H`` times;
    that needs ``model.check_finite = False`` (the reference's two sanity checks are host syncs).
    """

    def __init__(self, model, last_seq: torch.Tensor, H: int, x_mark: Optional[torch.Tensor] = None,
                 y_mark: Optional[torch.Tensor] = None, series_static: Optional[torch.Tensor] = None,
                 series_ids: Optional[torch.Tensor] = None, graph: bool = False):
        if not _is_native_model(model):
            raise TypeError("RecursiveForecaster drives the B200 TimesNet")
        if last_seq.ndim != 3:
            raise ValueError("last_seq must be shaped [B, T, N]")
        nv.require_cuda(last_seq, "last_seq")
        _check_marks(x_mark, y_mark, H)
        self.model, self.H = model, int(H)
        L = int(model.input_len)
        if last_seq.size(1) < L:
            raise ValueError(f"Input sequence length {last_seq.size(1)} is shorter than required input_len {L}")
        B, _, N = last_seq.shape
        dev = last_seq.device
        self.window = torch.empty(B, L, N, dtype=torch.float32, device=dev)
        self.mark = None if x_mark is None else torch.empty(B, L, x_mark.size(-1), dtype=torch.float32, device=dev)
        self.y_mark = None if x_mark is None else torch.empty(B, self.H, x_mark.size(-1), dtype=torch.float32, device=dev)
        self.rates = torch.empty(B, self.H, N, dtype=torch.float32, device=dev)
        self.disps = torch.empty(B, self.H, N, dtype=torch.float32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.static, self.ids = series_static, series_ids
        self._graph = None
        self._load(last_seq, x_mark, y_mark)
        if graph:
            if getattr(model, "check_finite", False):
                raise RuntimeError("graph replay needs model.check_finite = False (the checks are host syncs)")
            from .cuda_graphs import GraphedCallable
            keep = (self.window.clone(), None if self.mark is None else self.mark.clone())
            self._graph = GraphedCallable(lambda w: self._one_step(), [self.window], params_of=model)
            self._graph._static_in = [self.window]         # the step works in place on self.window: no input copy
            self.window.copy_(keep[0])
            if self.mark is not None:
                self.mark.copy_(keep[1])

    def _load(self, last_seq, x_mark, y_mark) -> None:
        L = self.window.size(1)
        self.window.copy_(last_seq[:, -L:, :])
        if self.mark is not None:
            self.mark.copy_(x_mark[:, -L:, :])
            self.y_mark.copy_(y_mark[:, : self.H, :])
        self.step.zero_()

    def _one_step(self) -> torch.Tensor:
        rate, disp = _invoke_model(self.model, self.window, x_mark=self.mark, series_static=self.static,
                                   series_ids=self.ids)
        if rate.size(1) != 1:
            raise RuntimeError("recursive forecasting needs a model built with mode='recursive' (one step per call)")
        nv.recursive_advance(self.window, rate, disp, self.rates, self.disps, self.mark, self.y_mark, self.step)
        return self.rates

    def run(self, last_seq: Optional[torch.Tensor] = None, x_mark: Optional[torch.Tensor] = None,
            y_mark: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``H`` rolling steps.  Returns the runner's own ``rates / disps`` buffers (clone to keep them)."""
        if last_seq is not None:
            self._load(last_seq, x_mark, y_mark)
        else:
            self.step.zero_()
        for _ in range(self.H):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._one_step()
        return self.rates, self.disps


def forecast_recursive_batch(model, last_seq: torch.Tensor, H: int, x_mark: Optional[torch.Tensor] = None,
                             y_mark: Optional[torch.Tensor] = None, series_static: Optional[torch.Tensor] = None,
                             series_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rolling one-step forecast: ``H`` full forwards, each appended to the window (predict.py:307-342)."""
    if _is_native_model(model) and last_seq.is_cuda and last_seq.dtype == torch.float32 and H > 0:
        runner = RecursiveForecaster(model, last_seq, H, x_mark, y_mark, series_static, series_ids)
        rates, disps = runner.run()
        return rates, disps
    # any other callable (a user wrapper around the model): the reference's host loop
    rates: List[torch.Tensor] = []
    disps: List[torch.Tensor] = []
    seq, mark_seq = last_seq, x_mark
    for step in range(H):
        rate_step, disp_step = _invoke_model(model, seq, x_mark=mark_seq, series_static=series_static,
                                             series_ids=series_ids)
        rates.append(rate_step)
        disps.append(disp_step)
        seq = torch.cat([seq[:, 1:, :], rate_step], dim=1)
        if mark_seq is not None:
            if y_mark is None:
                raise ValueError(
                    "Temporal features provided for history but missing future marks during recursive forecast")
            if y_mark.size(1) <= step:
                raise ValueError("y_mark does not provide enough future steps for recursive forecasting")
            mark_seq = torch.cat([mark_seq[:, 1:, :], y_mark[:, step:step + 1, :]], dim=1)
    return torch.cat(rates, dim=1), torch.cat(disps, dim=1)
