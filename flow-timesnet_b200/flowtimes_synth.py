"""Seeded synthetic inputs and weights for the BASELINE.json workloads.

Shared by ``bench.py``, ``tests/`` and ``oracle/make_golden.py`` so the CUDA
path, the oracle and the real reference all see identical tensors.  The input
generator follows the spec in SURVEY.md section 8(d): planted sinusoids with periods
24/12/7/48/6, per-(window, series) random phases, 0.1 sigma noise, offset 5.

Nothing here touches CUDA; tensors are created on the CPU with explicit
``torch.Generator`` objects and moved by the caller.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

PLANTED = ((24, 3.0), (12, 1.5), (7, 1.0), (48, 0.8), (6, 0.5))


@dataclass(frozen=True)
class Workload:
    name: str
    B: int
    T: int
    N: int
    H: int
    d_model: int
    n_layers: int
    k_periods: int
    dtype: str                    # "f32" | "bf16": activation dtype of the TimesBlock stack
    d_ff: int = 0                 # 0 -> 4 * d_model (config.py:183 default)
    kernel_set: Tuple[Tuple[int, int], ...] = ((3, 3), (5, 5), (7, 7))
    bottleneck_ratio: float = 4.0
    min_period_threshold: int = 1
    mode: str = "direct"
    context_rank: int = 0
    static_features: int = 0

    @property
    def ff(self) -> int:
        return self.d_ff if self.d_ff > 0 else 4 * self.d_model

    def as_dict(self) -> dict:
        d = asdict(self)
        d["d_ff"] = self.ff
        d["kernel_set"] = [list(k) for k in self.kernel_set]
        return d


# BASELINE.json configs[1..4] (configs[0] is the CSV pipeline run, out of scope).
WORKLOADS: Dict[str, Workload] = {
    "etth1": Workload("etth1", B=256, T=96, N=7, H=96, d_model=64, n_layers=2, k_periods=5, dtype="f32"),
    "elec": Workload("elec", B=64, T=336, N=321, H=96, d_model=128, n_layers=2, k_periods=5, dtype="bf16"),
    "traffic": Workload("traffic", B=32, T=720, N=862, H=336, d_model=256, n_layers=3, k_periods=5, dtype="f32"),
    "recursive": Workload("recursive", B=30000, T=28, N=1, H=28, d_model=128, n_layers=2, k_periods=2,
                          dtype="f32", d_ff=512, min_period_threshold=7, mode="recursive",
                          context_rank=16, static_features=5),
    # toy shapes used by the parity tests and the golden fixtures
    "toy": Workload("toy", B=4, T=48, N=5, H=12, d_model=16, n_layers=2, k_periods=3, dtype="f32", d_ff=32),
    "toy_bf16": Workload("toy_bf16", B=4, T=48, N=5, H=12, d_model=16, n_layers=2, k_periods=3, dtype="bf16",
                         d_ff=32),
    "mid": Workload("mid", B=8, T=96, N=7, H=24, d_model=64, n_layers=2, k_periods=5, dtype="f32"),
}


def planted_series(B: int, T: int, N: int, seed: int = 0) -> torch.Tensor:
    """``x[B,T,N]`` fp32 per SURVEY.md section 8(d) generator spec."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float32).view(1, T, 1)
    x = torch.zeros(B, T, N, dtype=torch.float32)
    for p, a in PLANTED:
        phi = torch.rand(B, 1, N, generator=g) * (2.0 * math.pi)
        x = x + a * torch.sin(2.0 * math.pi * t / float(p) + phi)
    x = x + 0.1 * torch.randn(B, T, N, generator=g)
    return x + 5.0


def planted_features(B: int, L: int, C: int, seed: int = 0) -> torch.Tensor:
    """Pre-embedded ``[B,L,d_model]`` features for the TimesBlock-stack scope.

    Same planted-period generator with channels in the role of series, centred
    (offset removed) and scaled to O(1) like an embedded, LayerNormed stream.
    """
    return (planted_series(B, L, C, seed) - 5.0) * 0.5


def white_features(B: int, L: int, C: int, seed: int = 1) -> torch.Tensor:
    """Adversarial white-noise features: several periods carry softmax weight."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, L, C, generator=g)


def poisson_targets(B: int, H: int, N: int, lam: float = 5.0, seed: int = 2) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.poisson(torch.full((B, H, N), lam), generator=g)


def _mid(cin: int, cout: int, ratio: float) -> int:
    return max(1, int(math.ceil(min(cin, cout) / float(ratio))))


def inception_shapes(prefix: str, cin: int, cout: int, kernel_set: Sequence[Sequence[int]],
                     ratio: float) -> Dict[str, Tuple[int, ...]]:
    """State-dict shapes of one InceptionBlock (reference timesnet.py:560-643)."""
    out: Dict[str, Tuple[int, ...]] = {}
    for j, (kh, kw) in enumerate(kernel_set):
        if math.isclose(ratio, 1.0, rel_tol=1e-9, abs_tol=1e-9):
            out[f"{prefix}paths.{j}.branch.0.weight"] = (cout, cin, kh, kw)
            out[f"{prefix}paths.{j}.branch.0.bias"] = (cout,)
        else:
            m = _mid(cin, cout, ratio)
            out[f"{prefix}paths.{j}.branch.0.weight"] = (m, cin, 1, 1)
            out[f"{prefix}paths.{j}.branch.0.bias"] = (m,)
            out[f"{prefix}paths.{j}.branch.1.weight"] = (m, m, kh, kw)
            out[f"{prefix}paths.{j}.branch.1.bias"] = (m,)
            out[f"{prefix}paths.{j}.branch.2.weight"] = (cout, m, 1, 1)
            out[f"{prefix}paths.{j}.branch.2.bias"] = (cout,)
    out[f"{prefix}proj.weight"] = (cout, cout * len(kernel_set), 1, 1)
    out[f"{prefix}proj.bias"] = (cout,)
    if cin != cout:
        out[f"{prefix}res_proj.weight"] = (cout, cin, 1, 1)
        out[f"{prefix}res_proj.bias"] = (cout,)
    return out


def stack_shapes(wl: Workload) -> Dict[str, Tuple[int, ...]]:
    shapes: Dict[str, Tuple[int, ...]] = {}
    for i in range(wl.n_layers):
        shapes.update(inception_shapes(f"blocks.{i}.inception.0.", wl.d_model, wl.ff, wl.kernel_set,
                                       wl.bottleneck_ratio))
        shapes.update(inception_shapes(f"blocks.{i}.inception.2.", wl.ff, wl.d_model, wl.kernel_set,
                                       wl.bottleneck_ratio))
    shapes["layer_norm.weight"] = (wl.d_model,)
    shapes["layer_norm.bias"] = (wl.d_model,)
    return shapes


def seeded_tensor(key: str, shape: Sequence[int], g: torch.Generator) -> torch.Tensor:
    """Deterministic value for one state-dict entry (drawn in sorted-key order)."""
    shape = tuple(int(s) for s in shape)
    r = torch.randn(shape, generator=g) if len(shape) > 0 else torch.randn((), generator=g)
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith("temporal_context.scale"):
        return torch.tensor(0.5)
    if key == "late_bias_gate":
        return torch.full(shape, 0.05) + 0.01 * r
    if key == "embedding.gate":
        return torch.full(shape, 0.1) + 0.02 * r
    if "norm" in key and leaf == "weight":
        return 1.0 + 0.1 * r
    if leaf == "bias":
        return 0.05 * r
    if key == "forecast_time_proj.weight":
        w = 0.5 * r / math.sqrt(shape[1])
        w[:, -1] += 1.0                         # keep the reference's copy-last-step prior
        return w
    if key == "series_embedding.weight":
        return r
    if key in ("mu_head.weight", "sigma_head.weight", "late_bias_head.weight", "context_coeff.weight",
               "context_proj.weight"):
        return 0.05 * r                          # randomise zero-init heads
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r / math.sqrt(fan_in)
    return r


def seeded_state(shapes: Dict[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(1000 + seed)
    return {k: seeded_tensor(k, shapes[k], g) for k in sorted(shapes)}


def stack_weights(wl: Workload, seed: int = 0) -> Dict[str, torch.Tensor]:
    return seeded_state(stack_shapes(wl), seed)


def reseed_module_state(module: torch.nn.Module, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded replacement for every entry of ``module.state_dict()`` (CPU fp32)."""
    sd = module.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items() if v is not None and v.dtype.is_floating_point}
    return seeded_state(shapes, seed)


def torch_dtype(name: str) -> torch.dtype:
    return {"f32": torch.float32, "bf16": torch.bfloat16}[name]


def stack_algorithmic_flops(wl: Workload, group_periods: Sequence[int]) -> float:
    """As-written conv FLOPs of one TimesBlock-stack forward."""
    C, Fh = wl.d_model, wl.ff
    m = _mid(C, Fh, wl.bottleneck_ratio)
    taps = sum(kh * kw for kh, kw in wl.kernel_set)
    nk = len(wl.kernel_set)
    if math.isclose(wl.bottleneck_ratio, 1.0):
        a = taps * C * Fh + nk * Fh * Fh + (C * Fh if C != Fh else 0)
        b = taps * Fh * C + nk * C * C + (C * Fh if C != Fh else 0)
    else:
        a = nk * C * m + taps * m * m + nk * m * Fh + nk * Fh * Fh + (C * Fh if C != Fh else 0)
        b = nk * Fh * m + taps * m * m + nk * m * C + nk * C * C + (C * Fh if C != Fh else 0)
    L = wl.T
    pos = sum(L + ((-L) % p) for p in group_periods)
    return 2.0 * (a + b) * pos * wl.B
