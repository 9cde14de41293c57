"""CUDA-graph replay of the forward path.

A TimesBlock forward is ~11 kernel launches and the period geometry never leaves the device
(``FtnPeriodPlan``), so a whole stack / model forward is capturable: no host round trip decides
anything between the first and the last kernel.  Replaying the captured graph removes the per-launch
host cost (Python + ctypes + driver), which at the BASELINE shapes is ~20 % of the step.

The reference trainer has an optional CUDA-graph capture of its step too (train.py:1261-1439).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple, Union

import torch

TensorOrTuple = Union[torch.Tensor, Tuple[torch.Tensor, ...]]


class GraphedCallable:
    """Capture ``fn(*inputs)`` once for fixed input shapes and replay it.

    ``fn`` must be free of host synchronisation and must enqueue all its work on the current stream
    (everything in ``timesnet_forecast`` does when ``check_finite`` is off).  Inputs are copied into
    static buffers before each replay; outputs are the static tensors of the capture (clone them if
    they have to survive the next call).
    """

    def __init__(self, fn: Callable[..., TensorOrTuple], example_inputs: Sequence[torch.Tensor], warmup: int = 2):
        if not example_inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedCallable needs CUDA tensors as example inputs (no CPU fallback)")
        self._fn = fn
        self._static_in = [t.detach().clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up: lazy builds, weight packing, smem attributes
            for _ in range(max(1, warmup)):
                fn(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_out = fn(*self._static_in)

    @property
    def inputs(self):
        """Static input tensors: write into them directly to skip the copy in ``__call__``."""
        return self._static_in

    def replay(self) -> TensorOrTuple:
        self._graph.replay()
        return self._static_out

    def __call__(self, *inputs: torch.Tensor) -> TensorOrTuple:
        if len(inputs) != len(self._static_in):
            raise ValueError(f"expected {len(self._static_in)} inputs, got {len(inputs)}")
        for dst, src in zip(self._static_in, inputs):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError("GraphedCallable was captured for inputs of shape "
                                 f"{tuple(dst.shape)}/{dst.dtype}, got {tuple(src.shape)}/{src.dtype}")
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._static_out
