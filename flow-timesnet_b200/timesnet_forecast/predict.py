"""Forecast helpers of the reference's predict.py that sit on the hot path
(predict.py:261-342).  The CSV / artifact pipeline around them is out of scope."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch


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


def forecast_recursive_batch(model, last_seq: torch.Tensor, H: int, x_mark: Optional[torch.Tensor] = None,
                             y_mark: Optional[torch.Tensor] = None, series_static: Optional[torch.Tensor] = None,
                             series_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rolling one-step forecast: H full forwards, each appended to the window (predict.py:307-342)."""
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
