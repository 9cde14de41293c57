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

    The capture bakes in the device pointers of the PACKED weight copies (``_pack.PackedInception``), not the
    ``nn.Parameter`` storage.  Pass the owning module as ``params_of``: every call then compares the parameters'
    (pointer, version) fingerprint with the one taken at capture time and re-captures after ``load_state_dict`` or an
    in-place update instead of silently replaying stale weights.
    """

    def __init__(self, fn: Callable[..., TensorOrTuple], example_inputs: Sequence[torch.Tensor], warmup: int = 2,
                 params_of: Optional[torch.nn.Module] = None):
        if not example_inputs or not all(isinstance(t, torch.Tensor) and t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedCallable needs CUDA tensors as example inputs (no CPU fallback)")
        self._fn = fn
        self._static_in = [t.detach().clone() for t in example_inputs]
        self._params_of = params_of
        self._warmup = warmup
        self.captures = 0
        self._capture()

    def _fingerprint(self):
        if self._params_of is None:
            return None
        return tuple((p.data_ptr(), p._version) for p in self._params_of.parameters())

    def _capture(self) -> None:
        fn, warmup = self._fn, self._warmup
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up: lazy builds, weight packing, smem attributes
            for _ in range(max(1, warmup)):
                fn(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        # capture on the warm-up stream: per-(device, stream) state the modules created while warming up (the zeroed
        # plan buffers of the TimesBlocks) is found again during the capture instead of being re-created inside it
        with torch.cuda.graph(self._graph, stream=side):
            self._static_out = fn(*self._static_in)
        self._captured_fp = self._fingerprint()
        self.captures += 1

    def _refresh(self) -> None:
        if self._params_of is not None and self._fingerprint() != self._captured_fp:
            torch.cuda.synchronize()
            self._capture()                                 # parameters changed: re-pack and re-capture

    @property
    def inputs(self):
        """Static input tensors: write into them directly to skip the copy in ``__call__``."""
        return self._static_in

    def replay(self) -> TensorOrTuple:
        self._refresh()
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
        self._refresh()
        self._graph.replay()
        return self._static_out


class PipelinedRunner:
    """Host-fed streaming of ``fn``: the host->device copy of batch ``i+1`` overlaps the graph replay of
    batch ``i``.

    Two captured graphs with their own static input buffers alternate; a dedicated copy stream fills
    the idle one from (pinned) host memory while the compute stream replays the other.  ``submit``
    enqueues copy + replay + device->host read-back of the result and returns immediately;
    ``results`` are pinned host tensors, valid after ``synchronize()`` (or after the matching
    event).  This is the serving loop a caller with host-resident windows uses
    (reference: predict.py:930-950 moves each batch to the device inside its forecast loop).
    """

    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor], depth: int = 2,
                 params_of: Optional[torch.nn.Module] = None):
        dev = example_inputs[0].device
        self._graphs = [GraphedCallable(fn, example_inputs, params_of=params_of) for _ in range(depth)]
        self._copy = torch.cuda.Stream(device=dev)
        self._back = torch.cuda.Stream(device=dev)                     # result read-back: off the compute stream, so the
                                                                       # next replay does not queue behind a DMA round trip
        self._filled = [torch.cuda.Event() for _ in range(depth)]      # inputs of slot s are on the device
        self._consumed = [torch.cuda.Event() for _ in range(depth)]    # slot s's replay has read its inputs
        self._read = [torch.cuda.Event() for _ in range(depth)]        # slot s's result is in its host buffer
        out = self._graphs[0]._static_out
        if not isinstance(out, torch.Tensor):
            raise TypeError("PipelinedRunner expects fn to return one tensor (e.g. the loss)")
        self.results = [torch.empty(out.shape, dtype=out.dtype).pin_memory() for _ in range(depth)]
        self._i = 0
        cur = torch.cuda.current_stream(dev)
        for e in self._consumed + self._read:
            e.record(cur)

    def submit(self, *host_inputs: torch.Tensor) -> int:
        """Enqueue one batch (host tensors, ideally pinned).  Returns the slot whose ``results`` entry it fills."""
        s = self._i % len(self._graphs)
        self._i += 1
        g = self._graphs[s]
        cur = torch.cuda.current_stream()
        self._copy.wait_event(self._consumed[s])            # the previous replay of this slot is done with the buffers
        with torch.cuda.stream(self._copy):
            for dst, src in zip(g.inputs, host_inputs):
                dst.copy_(src, non_blocking=True)
            self._filled[s].record(self._copy)
        cur.wait_event(self._filled[s])
        cur.wait_event(self._read[s])                       # the slot's previous result has left its (static) output tensor
        out = g.replay()
        self._consumed[s].record(cur)
        self._back.wait_event(self._consumed[s])
        with torch.cuda.stream(self._back):
            self.results[s].copy_(out, non_blocking=True)
            self._read[s].record(self._back)
        return s

    def synchronize(self) -> None:
        torch.cuda.current_stream().synchronize()
        self._back.synchronize()
