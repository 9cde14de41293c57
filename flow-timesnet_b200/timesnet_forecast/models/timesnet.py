"""B200-native drop-in for ``timesnet_forecast.models.timesnet`` (forward path).

Same public surface as the reference module (class names, constructor
signatures, ``forward`` signatures, ``state_dict`` key names, diagnostic
attributes -- SURVEY.md section 8b) but every tensor op of the TimesBlock path runs in
hand-written sm_100a CUDA kernels reached through the C ABI of
``libflowtimes.so`` (include/flowtimes.h).  PyTorch only provides parameters,
device memory and streams.

Differences a caller can observe, all deliberate:
  * CUDA only, float32 / bfloat16 activations only -- CPU tensors and fp16 raise
    (BASELINE.json north_star: "no CPU fallback").
  * forward-only: outputs carry no autograd graph, dropout is the identity
    (inference semantics; the reference's eval mode).  A module left in train
    mode with dropout > 0 warns once; an input that requires grad raises.
  * multi-rank period search is OPT-IN: like the reference under DDP every rank
    selects periods from its local batch unless
    ``timesnet_forecast.parallel.share_period_search(model)`` was called.
  * periods stay on the device: ``TimesBlock`` no longer does the reference's
    ~20 ``.item()`` syncs per call.  The diagnostic attributes
    (``last_selected_periods``, ``_last_group_count`` ...) read the plan back
    lazily, i.e. they sync only when somebody looks at them.
  * the performance-tuning env flags ``TIMESBLOCK_MEMORY_FORMAT``,
    ``TIMESBLOCK_VEC_DISABLE``, ``TIMESBLOCK_BUCKET_MAX``, ``TIMESBLOCK_K_CHUNK``,
    ``TIMES_MP_CONV`` are accepted and ignored (they select between numerically
    identical reference code paths).  ``TIMES_PERIOD_BINNING`` /
    ``TIMES_PERIOD_MAX_UNIQ`` change results and are honoured.

Reference citations are ``timesnet.py:<line>`` of
/root/reference/src/timesnet_forecast/models/timesnet.py.
"""
from __future__ import annotations

import math
import os
import warnings
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from .. import _native as nv
from .._pack import PackedInception, pack_inception_block, params_fingerprint, split_linear_weight
from ..parallel import reduce_spectrum_sum, resolve_group

__all__ = [
    "FFTPeriodSelector", "PeriodGroupResult", "PeriodGrouper", "InceptionBranch", "InceptionBlock", "TimesBlock",
    "PositionalEmbedding", "RMSNorm", "DataEmbedding", "LowRankTemporalContext", "TimesNet",
]


def _act_code(name: str) -> int:
    return nv.FTN_ACT_RELU if name == "relu" else nv.FTN_ACT_GELU


_warned_train = set()


def _wants_grad(module: nn.Module, x: torch.Tensor) -> bool:
    """The differentiable route of a module is taken when autograd is recording and either the input carries a graph or
    the module was opted in (``module.differentiable = True``: weight gradients for a grad-less input)."""
    return torch.is_grad_enabled() and (bool(getattr(x, "requires_grad", False)) or bool(getattr(module, "differentiable", False)))


def _forward_only_guard(module: nn.Module, x: torch.Tensor, dropout: float) -> None:
    """The build is forward-only: say so instead of silently returning grad-less, dropout-free outputs."""
    if torch.is_grad_enabled() and isinstance(x, torch.Tensor) and x.requires_grad:
        raise RuntimeError(
            f"{type(module).__name__}: input requires grad, but the B200 build is forward-only (its outputs carry no "
            "autograd graph); detach the input or wrap the call in torch.no_grad()")
    if module.training and dropout > 0.0 and id(module) not in _warned_train:
        _warned_train.add(id(module))
        warnings.warn(
            f"{type(module).__name__} is in train mode with dropout={dropout}: the B200 build is forward-only and runs "
            "the eval-mode path (dropout = identity); call .eval()", RuntimeWarning, stacklevel=3)


# --------------------------------------------------------------------------- #
# device-resident period plan
# --------------------------------------------------------------------------- #
class PeriodPlan:
    """Result of one period search: device plan + per-window amplitudes/weights.

    ``host()`` copies the 1 KB plan struct back (one sync) and is only used by
    the diagnostic attributes and the public ``FFTPeriodSelector.forward``.
    """

    def __init__(self, plan_dev: torch.Tensor, amps: Optional[torch.Tensor], weights: torch.Tensor, k: int):
        self.plan_dev = plan_dev
        self.amps = amps
        self.weights = weights
        self.k = k
        self._host: Optional[nv.FtnPeriodPlan] = None

    def host(self) -> nv.FtnPeriodPlan:
        if self._host is None:
            self._host = nv.plan_to_host(self.plan_dev)
        return self._host


class FFTPeriodSelector(nn.Module):
    """Shared dominant-period selector (reference timesnet.py:52-159).

    ``forward`` keeps the reference contract ``x[B,L,C] -> (periods[K] long,
    amplitudes[B,K] x.dtype)``.  ``search`` is the sync-free entry TimesBlock
    uses.  Like the reference under DDP, every rank selects periods from its
    LOCAL batch by default (``process_group = False``).  Opt in to the shared
    search of SURVEY.md section 8e with ``parallel.share_period_search(model, group)``
    (sets ``process_group`` to a group, or ``None`` for the default group): the
    batch-summed spectrum (L/2+1 floats + the window count) is then all-reduced
    so every rank selects the same periods.  Every rank must call ``forward``
    the same number of times; a rank holding zero windows still enters the
    collective.
    """

    def __init__(self, k_periods: int, pmax: int, min_period_threshold: int = 1) -> None:
        super().__init__()
        self.k = int(max(0, k_periods))
        self.pmax = int(max(1, pmax))
        self.min_period_threshold = int(min(self.pmax, int(max(1, min_period_threshold))))
        self.process_group = False         # False = rank-local (reference behaviour); None / a group = shared search
        self.peer_comm = None              # nv.PeerComm: exchange over NVLink peer memory inside the selection kernel
        self._last_plan: Optional[PeriodPlan] = None
        self._empty_device = torch.device("cpu")

    # -- lazily materialised diagnostics (timesnet.py:156-157) ------------------
    @property
    def last_frequency_indices(self) -> torch.Tensor:
        if self._last_plan is None:
            return torch.zeros(0, dtype=torch.long, device=self._empty_device)
        h = self._last_plan.host()
        return torch.tensor(list(h.freq[: h.n_valid]), dtype=torch.long, device=self._last_plan.plan_dev.device)

    @property
    def last_selected_periods(self) -> torch.Tensor:
        if self._last_plan is None:
            return torch.zeros(0, dtype=torch.long, device=self._empty_device)
        h = self._last_plan.host()
        return torch.tensor(list(h.period[: h.n_valid]), dtype=torch.long, device=self._last_plan.plan_dev.device)

    def _world(self):
        return resolve_group(self.process_group)

    def search(self, x: torch.Tensor) -> Optional[PeriodPlan]:
        """Sync-free search.  Returns None when the reference would return empty tensors."""
        if x.ndim != 3:
            raise ValueError("FFTPeriodSelector expects input shaped [B, L, C]")
        B, L, C = x.shape
        self._last_plan = None
        self._empty_device = x.device
        if self.k <= 0 or L <= 1 or C <= 0:
            return None                                              # timesnet.py:89-90 (same on every rank)
        nbins = L // 2 + 1
        k = min(self.k, nbins - 1)                                   # timesnet.py:122-126
        if k <= 0:
            return None
        if k > nv.FTN_MAX_K:
            raise ValueError(f"k_periods={self.k} exceeds the supported maximum {nv.FTN_MAX_K}")
        group, world = self._world()
        if B <= 0:
            if world > 1:                                            # an empty shard still enters the collective
                zeros = torch.zeros(nbins + 1, dtype=torch.float32, device=x.device)
                if self.peer_comm is not None:
                    self.peer_comm.all_reduce(zeros)
                else:
                    reduce_spectrum_sum(zeros, self.process_group)
            return None                                              # timesnet.py:89-90
        x = nv.require_cuda(x, "x")
        if world == 1 or self.peer_comm is not None:                   # fused 2-launch search (peer exchange in-kernel)
            plan_dev, amps, weights, _, _ = nv.period_search(x, k, self.pmax, self.min_period_threshold,
                                                             self.peer_comm if world > 1 else None)
            self._last_plan = PeriodPlan(plan_dev, amps, weights, k)
            return self._last_plan
        med, ssum = nv.spectrum(x)                                     # ssum = [sum_b median spectrum | B]
        # the only collective of the path: F sums + the window count in one message, so ragged
        # shards need no host round trip (the select kernel divides by the reduced count)
        reduce_spectrum_sum(ssum, self.process_group)
        plan_dev, amps, weights = nv.select_periods(med, ssum, x.dtype, 0, L, k, self.pmax, self.min_period_threshold)
        self._last_plan = PeriodPlan(plan_dev, amps, weights, k)
        return self._last_plan

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if x.ndim != 3:
            raise ValueError("FFTPeriodSelector expects input shaped [B, L, C]")
        B = x.shape[0]
        plan = self.search(x)
        if plan is None:
            return (torch.zeros(0, dtype=torch.long, device=x.device),
                    torch.zeros(B, 0, dtype=x.dtype, device=x.device))
        h = plan.host()
        nvld = h.n_valid
        periods = torch.tensor(list(h.period[:nvld]), dtype=torch.long, device=x.device)
        return periods, plan.amps[:, :nvld].contiguous()


# --------------------------------------------------------------------------- #
# PeriodGrouper (host-side integer logic, K <= 16 scalars)
# --------------------------------------------------------------------------- #
def _scheduled_token(raw: Optional[str], depth: Optional[int]) -> Optional[str]:
    """Depth-scheduled env syntax ``"0:3,2:2,default:4"`` (timesnet.py:162-216)."""
    if raw is None or not raw.strip():
        return None
    toks = [t.strip() for t in raw.split(",") if t.strip()]
    if not toks:
        return None
    plain: List[str] = []
    default: List[str] = []
    keyed = {}
    for tok in toks:
        sep = ":" if ":" in tok else ("=" if "=" in tok else None)
        if sep is None:
            plain.append(tok)
            continue
        key, val = (s.strip() for s in tok.split(sep, 1))
        if not val:
            continue
        if key.lower() in ("default", "*"):
            default.append(val)
        else:
            try:
                keyed[int(key)] = val
            except ValueError:
                pass
    if depth is not None and keyed:
        if depth in keyed:
            return keyed[depth]
        below = [d for d in keyed if d <= depth]
        if below:
            return keyed[max(below)]
    if default:
        return default[-1]
    if plain:
        return plain[-1]
    if keyed:
        return keyed[min(keyed)]
    return toks[-1]


def _resolve_scheduled_int(raw: Optional[str], depth: Optional[int]) -> Optional[int]:
    tok = _scheduled_token(raw, depth)
    if tok is None:
        return None
    try:
        val = int(float(tok))
    except ValueError:
        return None
    return val if val > 0 else None


def _resolve_log_binning_base(raw: Optional[str], depth: Optional[int]) -> Optional[float]:
    tok = _scheduled_token(raw, depth)
    if tok is None:
        return None
    text = tok.strip().lower()
    if not text or text in ("off", "false", "0", "none"):
        return None
    names = ("log", "logscale", "logarithmic")
    base: Optional[float] = None
    if ":" in text:
        head, tail = (s.strip() for s in text.split(":", 1))
        try:
            base = float(tail) if head in names else float(head)
        except ValueError:
            base = None
    elif text in names:
        base = 2.0
    else:
        try:
            base = float(text)
        except ValueError:
            base = None
    if base is None:
        base = 2.0
    return float(base) if base > 1.0 else None


@dataclass
class PeriodGroupResult:
    periods: torch.Tensor
    pad_lengths: torch.Tensor
    cycles: torch.Tensor
    logits: torch.Tensor
    mapping: torch.Tensor
    valid_mask: torch.Tensor
    canonical_indices: torch.Tensor


class PeriodGrouper:
    """Duplicate / binning-aware grouping of candidate periods (timesnet.py:286-557).

    Integer work on at most 16 scalars: done on the host with Python ints; the
    amplitudes are only touched for the per-group ``logsumexp`` logits and the
    canonical-member choice.  The CUDA path uses this class only for custom
    selector modules or when the ``TIMES_PERIOD_*`` opt-ins are set; the default
    grouping of the fused path runs on the device (csrc/common.cuh).
    """

    def __init__(self, periods: torch.Tensor, amplitudes: torch.Tensor, seq_len: int, *,
                 min_period: Optional[int] = None, max_period: Optional[int] = None,
                 block_index: Optional[int] = None, freq_indices: Optional[torch.Tensor] = None) -> None:
        self.periods = periods.view(-1)
        amp = amplitudes.unsqueeze(0) if amplitudes.dim() == 1 else amplitudes
        if amp.dim() != 2:
            raise ValueError("amplitudes must have shape [B, K] or [K]")
        if amp.size(1) != self.periods.numel():
            raise ValueError("amplitudes second dimension must match number of period candidates")
        self.amplitudes = amp
        self.seq_len = int(seq_len)
        self.device = self.periods.device
        self.batch = amp.size(0)
        self.amp_dtype = amp.dtype
        self.period_dtype = self.periods.dtype
        self.min_period = None if min_period is None else int(min_period)
        self.max_period = None if max_period is None else int(max_period)
        self.block_index = None if block_index is None else int(block_index)
        self.freq_indices = freq_indices
        self.max_unique = _resolve_scheduled_int(os.getenv("TIMES_PERIOD_MAX_UNIQ"), self.block_index)
        self.log_base = _resolve_log_binning_base(os.getenv("TIMES_PERIOD_BINNING"), self.block_index)

    def _empty(self) -> PeriodGroupResult:
        K = self.periods.numel()
        zl = torch.zeros(0, dtype=self.period_dtype, device=self.device)
        return PeriodGroupResult(
            periods=zl, pad_lengths=zl, cycles=zl,
            logits=torch.zeros(self.batch, 0, dtype=self.amp_dtype, device=self.amplitudes.device),
            mapping=torch.full((K,), -1, dtype=torch.long, device=self.device),
            valid_mask=torch.zeros(K, dtype=torch.bool, device=self.device),
            canonical_indices=torch.zeros(0, dtype=torch.long, device=self.device))

    def _bucket(self, p: int) -> int:
        val = torch.log(torch.tensor(float(p), dtype=torch.float32)) / math.log(self.log_base)
        return int(torch.floor(val + 1e-6).item())

    def group(self) -> PeriodGroupResult:
        L = self.seq_len
        cand: List[Tuple[int, int, int, int]] = []            # (candidate index, period, pad, cycles)
        for i, p in enumerate(self.periods.tolist()):
            if p <= 0:
                continue
            if self.min_period is not None and p < self.min_period:
                continue
            if self.max_period is not None and p > self.max_period:
                continue
            pad = (-L) % p
            cyc = (L + pad) // p
            if cyc >= 2:
                cand.append((i, p, pad, cyc))
        if not cand:
            return self._empty()
        amp_sel = self.amplitudes[:, [c[0] for c in cand]]
        keys = [self._bucket(c[1]) if self.log_base is not None else c[1] for c in cand]
        order = sorted(set(keys))
        assign = [order.index(v) for v in keys]

        def describe(assign_now: List[int]):
            groups = []
            for gid in sorted(set(assign_now)):
                members = [j for j, a in enumerate(assign_now) if a == gid]
                cols = amp_sel[:, members]
                logits = torch.logsumexp(cols, dim=1)
                lead = 0 if len(members) == 1 else int(torch.argmax(cols.mean(dim=0)).item())
                canon = members[lead]
                groups.append(dict(id=gid, members=members, canon=canon, logits=logits,
                                   score=float(logits.mean().item())))
            return groups

        if self.max_unique is not None and len(set(assign)) > self.max_unique:
            groups = describe(assign)
            scores = torch.tensor([g["score"] for g in groups], dtype=torch.float32)
            keep = torch.topk(scores, k=self.max_unique, largest=True).indices.tolist()
            keep_p = torch.tensor([float(cand[groups[j]["canon"]][1]) for j in keep], dtype=torch.float32)
            merged = list(assign)
            for j, g in enumerate(groups):
                if j in keep:
                    continue
                near = int(torch.argmin(torch.abs(keep_p - float(cand[g["canon"]][1]))).item())
                for m in g["members"]:
                    merged[m] = groups[keep[near]]["id"]
            assign = merged

        groups = describe(assign)
        groups.sort(key=lambda g: (cand[g["canon"]][1], cand[g["canon"]][0]))
        K = self.periods.numel()
        mapping = [-1] * K
        valid = [False] * K
        for c in cand:
            valid[c[0]] = True
        for gi, g in enumerate(groups):
            for m in g["members"]:
                mapping[cand[m][0]] = gi
        mk = lambda vals, dt: torch.tensor(vals, dtype=dt, device=self.device)
        return PeriodGroupResult(
            periods=mk([cand[g["canon"]][1] for g in groups], self.period_dtype),
            pad_lengths=mk([cand[g["canon"]][2] for g in groups], self.period_dtype),
            cycles=mk([cand[g["canon"]][3] for g in groups], self.period_dtype),
            logits=torch.stack([g["logits"] for g in groups], dim=1),
            mapping=mk(mapping, torch.long),
            valid_mask=mk(valid, torch.bool),
            canonical_indices=mk([cand[g["canon"]][0] for g in groups], torch.long))


def _plan_from_group_result(res: PeriodGroupResult, n_candidates: int, L: int) -> nv.FtnPeriodPlan:
    """Fill the C plan struct from an explicit grouping (binning / max-uniq modes)."""
    pl = nv.FtnPeriodPlan()
    G = int(res.periods.numel())
    if G > nv.FTN_MAX_K or n_candidates > nv.FTN_MAX_K:
        raise ValueError(f"at most {nv.FTN_MAX_K} candidate periods are supported")
    pl.seq_len, pl.n_raw, pl.n_valid, pl.n_groups = L, n_candidates, n_candidates, G
    mapping = res.mapping.tolist()
    for i in range(nv.FTN_MAX_K):
        pl.mapping[i] = mapping[i] if i < n_candidates else -1
    off = 0
    per, pad, cyc, canon = res.periods.tolist(), res.pad_lengths.tolist(), res.cycles.tolist(), \
        res.canonical_indices.tolist()
    for g in range(nv.FTN_MAX_K):
        pl.grp_row_off[g] = off
        if g < G:
            pl.grp_period[g], pl.grp_pad[g], pl.grp_cycles[g], pl.grp_canon[g] = per[g], pad[g], cyc[g], canon[g]
            off += L + pad[g]
        else:
            pl.grp_canon[g] = -1
    pl.grp_row_off[nv.FTN_MAX_K] = off
    pl.total_rows_per_window = off
    return pl


# --------------------------------------------------------------------------- #
# Inception bank (parameter containers; compute goes through libflowtimes)
# --------------------------------------------------------------------------- #
class InceptionBranch(nn.Module):
    """One branch: k x k conv, or 1x1 -> k x k -> 1x1 bottleneck (timesnet.py:560-593)."""

    def __init__(self, in_ch: int, out_ch: int, kernel_size: Tuple[int, int], bottleneck_ratio: float) -> None:
        super().__init__()
        if bottleneck_ratio <= 0:
            raise ValueError("bottleneck_ratio must be a positive value")
        kh, kw = kernel_size
        pad = (max(kh // 2, 0), max(kw // 2, 0))
        if math.isclose(bottleneck_ratio, 1.0, rel_tol=1e-9, abs_tol=1e-9):
            layers = [nn.Conv2d(in_ch, out_ch, kernel_size=(kh, kw), padding=pad)]
        else:
            mid = max(1, int(math.ceil(min(in_ch, out_ch) / float(bottleneck_ratio))))
            layers = [nn.Conv2d(in_ch, mid, kernel_size=1),
                      nn.Conv2d(mid, mid, kernel_size=(kh, kw), padding=pad),
                      nn.Conv2d(mid, out_ch, kernel_size=1)]
        self.branch = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``branch(x)`` on an NCHW grid (timesnet.py:592-593): each Conv2d runs as an implicit GEMM on the
        zero-copy fold of the grid (``ftn_conv2d_grid``; one group, period = W, H cycles)."""
        if x.ndim != 4:
            raise ValueError("InceptionBranch expects an NCHW grid [B, C, H, W]")
        x = nv.require_cuda(x, "x")
        if _wants_grad(self, x):                                   # differentiable route: native forward AND backward
            from ..autograd import inception_branch
            return inception_branch(x, self.branch).to(x.dtype)
        B, _, H, W = x.shape
        dt = x.dtype
        with torch.no_grad():
            seq = x.permute(0, 2, 3, 1).reshape(B, H * W, -1).to(torch.float32).contiguous()
            plan = nv.single_group_plan(W, H, x.device)
            for conv in self.branch:
                kh, kw = conv.kernel_size
                if kh % 2 == 0 or kw % 2 == 0:
                    raise ValueError("libflowtimes convolutions need odd kernel sizes (\"same\" padding k // 2)")
                w = conv.weight.detach().to(device=x.device, dtype=torch.float32)
                taps = w.permute(2, 3, 1, 0).reshape(kh * kw, w.shape[1], w.shape[0]).contiguous()
                seq = nv.conv2d_grid(seq, plan, taps, conv.bias.detach().to(device=x.device, dtype=torch.float32).contiguous(),
                                     kh, kw)
            return seq.reshape(B, H, W, -1).permute(0, 3, 1, 2).to(dt)


def _parse_kernel_set(kernel_set) -> List[Tuple[int, int]]:
    spec: List[Tuple[int, int]] = []
    for k in kernel_set:
        if isinstance(k, tuple):
            kh, kw = k
        elif isinstance(k, Sequence):
            if len(k) != 2:
                raise ValueError("kernel_set entries must be (kh, kw) pairs")
            kh, kw = k
        else:
            kh = kw = int(k)
        spec.append((int(kh), int(kw)))
    if not spec:
        raise ValueError("kernel_set must contain at least one kernel size")
    return spec


class InceptionBlock(nn.Module):
    """Inception block on the cycle/period grid (timesnet.py:596-654).

    Sub-module names (``paths[i].branch[j]``, ``proj``, ``res_proj``) and
    parameter shapes match the reference so checkpoints load unchanged.
    ``forward`` accepts an NCHW grid like the reference and runs the packed
    block on the device by treating the grid as one period group (period = W).
    """

    def __init__(self, in_ch: int, out_ch: int, kernel_set, dropout: float, act: str,
                 bottleneck_ratio: float = 1.0) -> None:
        super().__init__()
        spec = _parse_kernel_set(kernel_set)
        self.paths = nn.ModuleList(
            [InceptionBranch(in_ch, out_ch, (kh, kw), bottleneck_ratio) for kh, kw in spec])
        self.proj = nn.Conv2d(out_ch * len(spec), out_ch, kernel_size=1)
        self.res_proj = nn.Conv2d(in_ch, out_ch, kernel_size=1) if in_ch != out_ch else nn.Identity()
        self.dropout = nn.Dropout(dropout)
        self.act = nn.ReLU() if act.lower() == "relu" else nn.GELU()
        self._packed: Optional[PackedInception] = None
        self._packed_key = None

    def packed(self, device: torch.device) -> PackedInception:
        key = (params_fingerprint(self), str(device))
        if self._packed is None or self._packed_key != key:
            self._packed = pack_inception_block(self, device)
            self._packed_key = key
        return self._packed

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``act(proj(cat_j paths_j(x))) + res_proj(x)`` on an NCHW grid (timesnet.py:645-654) through the packed
        block (``ftn_inception_block``): the grid is one period group of the zero-copy fold (period W, H cycles)."""
        if x.ndim != 4:
            raise ValueError("InceptionBlock expects an NCHW grid [B, C, H, W]")
        x = nv.require_cuda(x, "x")
        if _wants_grad(self, x):
            # differentiable route (second backward slice): every conv / activation is a libflowtimes kernel forward
            # and backward (timesnet_forecast/autograd.py); as written, fp32, not the packed tensor-core chain
            if self.training and float(self.dropout.p) > 0.0:
                raise RuntimeError("InceptionBlock: the differentiable route has no dropout (set dropout=0 or call .eval())")
            from ..autograd import inception_block
            if self.proj.weight.device != x.device:
                self.to(x.device)
            return inception_block(x, self).to(x.dtype)
        _forward_only_guard(self, x, float(self.dropout.p))
        B, _, H, W = x.shape
        dt = x.dtype
        if self.proj.weight.device != x.device:
            self.to(x.device)
        with torch.no_grad():
            seq = x.permute(0, 2, 3, 1).reshape(B, H * W, -1).to(torch.float32).contiguous()
            act = nv.FTN_ACT_RELU if isinstance(self.act, nn.ReLU) else nv.FTN_ACT_GELU
            out = nv.inception_block(seq, nv.single_group_plan(W, H, x.device), self.packed(x.device).struct, act, False)
            return out.reshape(B, H, W, -1).permute(0, 3, 1, 2).to(dt)


# --------------------------------------------------------------------------- #
# TimesBlock
# --------------------------------------------------------------------------- #
class TimesBlock(nn.Module):
    """TimesNet block on ``[B, L, C]`` features (timesnet.py:657-1101)."""

    def __init__(self, d_model: Optional[int], kernel_set, dropout: float, activation: str,
                 d_ff: Optional[int] = None, bottleneck_ratio: float = 1.0) -> None:
        super().__init__()
        self._configured_d_model = int(d_model) if d_model is not None else None
        if d_ff is None:
            self._configured_d_ff: Optional[int] = None
        else:
            self._configured_d_ff = int(d_ff)
            if self._configured_d_ff <= 0:
                raise ValueError("d_ff must be a positive integer")
        self.d_model: Optional[int] = None
        self.d_ff: Optional[int] = None
        self.bottleneck_ratio = float(bottleneck_ratio)
        if self.bottleneck_ratio <= 0:
            raise ValueError("bottleneck_ratio must be a positive value")
        self._activation_name = "relu" if activation.lower() == "relu" else "gelu"
        self._kernel_spec = _parse_kernel_set(kernel_set)
        self._dropout = float(dropout)
        self.inception: Optional[nn.Module] = None
        if self._configured_d_model is not None:
            self._build_layers(self._configured_d_model, device=torch.device("cpu"), dtype=torch.get_default_dtype())
        self.period_selector = None        # injected by TimesNet via object.__setattr__ (timesnet.py:711-713)
        self._period_calls: int = 0
        self._vec_calls: int = 0
        self.block_index: Optional[int] = None
        self._last_plan: Optional[PeriodPlan] = None
        self._last_raw: int = 0

    # -- diagnostics the reference tests read (timesnet.py:714-720); lazy = sync on access --
    @property
    def _last_raw_period_count(self) -> int:
        if self._last_raw >= 0:
            return self._last_raw
        return 0 if self._last_plan is None else int(self._last_plan.host().n_valid)

    @property
    def _last_valid_period_count(self) -> int:
        if self._last_plan is None:
            return 0
        return sum(1 for m in self._last_plan.host().mapping[: nv.FTN_MAX_K] if m >= 0)

    @property
    def _last_group_count(self) -> int:
        return 0 if self._last_plan is None else int(self._last_plan.host().n_groups)

    @property
    def _last_loop_iterations(self) -> int:
        return self._last_group_count

    def _build_layers(self, channels: int, device: torch.device, dtype: torch.dtype) -> None:
        if channels <= 0:
            raise ValueError("TimesBlock requires a positive channel count")
        self.d_model = int(channels)
        hidden = self._configured_d_ff if self._configured_d_ff is not None else self.d_model
        if hidden <= 0:
            raise ValueError("Derived hidden dimension must be positive")
        self.d_ff = int(hidden)
        mid_act: nn.Module = nn.ReLU() if self._activation_name == "relu" else nn.GELU()
        self.inception = nn.Sequential(
            InceptionBlock(self.d_model, self.d_ff, self._kernel_spec, self._dropout, self._activation_name,
                           self.bottleneck_ratio),
            mid_act,
            InceptionBlock(self.d_ff, self.d_model, self._kernel_spec, self._dropout, self._activation_name,
                           self.bottleneck_ratio),
        ).to(device=device, dtype=torch.float32)      # conv weights stay fp32 (timesnet.py:14-34)

    def _is_native_bank(self) -> bool:
        inc = self.inception
        return (isinstance(inc, nn.Sequential) and len(inc) == 3 and isinstance(inc[0], InceptionBlock)
                and isinstance(inc[2], InceptionBlock))

    # ------------------------------------------------------------------ #
    def _plan_for(self, x: torch.Tensor) -> Optional[PeriodPlan]:
        """Run the selector and return a device plan (None = identity block)."""
        sel = self.period_selector
        B, L, C = x.shape
        env_grouping = bool(os.getenv("TIMES_PERIOD_MAX_UNIQ") or os.getenv("TIMES_PERIOD_BINNING"))
        self._last_plan = None
        if isinstance(sel, FFTPeriodSelector) and not env_grouping:
            plan = sel.search(x)                                # sync-free fused path
            self._last_raw = 0 if plan is None else -1          # -1: read n_valid from the plan on demand
            self._last_plan = plan
            return plan
        periods, amplitudes = sel(x)                            # arbitrary selector module (host-visible)
        periods = periods.reshape(-1).to(torch.long)
        self._last_raw = int(periods.numel())
        if periods.numel() == 0:
            return None                                         # timesnet.py:796-797
        amplitudes = amplitudes.to(device=x.device, dtype=x.dtype)
        if periods.numel() > nv.FTN_MAX_K:
            raise ValueError(f"at most {nv.FTN_MAX_K} candidate periods are supported")
        min_p = getattr(sel, "min_period_threshold", None)
        max_p = getattr(sel, "pmax", None)
        per_host = periods.tolist()
        if env_grouping:
            res = PeriodGrouper(periods.cpu(), amplitudes.float().cpu(), L, min_period=min_p, max_period=max_p,
                                block_index=self.block_index).group()
            host_plan = _plan_from_group_result(res, len(per_host), L)
        else:
            host_plan = nv.plan_build_host(per_host, L, min_p, max_p)
        for i, p in enumerate(per_host):
            host_plan.period[i] = int(p)
        plan_dev = nv.plan_to_device(host_plan, x.device)
        if host_plan.n_groups == 0:
            self._last_plan = PeriodPlan(plan_dev, None, None, len(per_host))
            self._last_plan._host = host_plan
            return None                                         # timesnet.py:989-990
        amps = nv.require_cuda(amplitudes, "amplitudes")
        weights = nv.group_weights(amps, plan_dev, B)
        plan = PeriodPlan(plan_dev, amps, weights, len(per_host))
        plan._host = host_plan
        self._last_plan = plan
        return plan

    def _custom_bank_delta(self, x: torch.Tensor, plan: PeriodPlan) -> torch.Tensor:
        """User-supplied ``inception`` module (reference tests swap one in): fold with
        views, call the module, subtract the grid.  The module is the compute here."""
        B, L, C = x.shape
        h = plan.host()
        delta = torch.empty(max(1, h.n_groups), B, L, C, dtype=x.dtype, device=x.device)
        xp = x.permute(0, 2, 1)
        for g in range(h.n_groups):
            p, pad, cyc = h.grp_period[g], h.grp_pad[g], h.grp_cycles[g]
            grid = torch.nn.functional.pad(xp, (0, pad)).reshape(B, C, cyc, p).to(torch.float32)
            out = self.inception(grid)
            d = (out.to(torch.float32) - grid).reshape(B, C, cyc * p)[..., :L]
            delta[g] = d.permute(0, 2, 1).to(x.dtype)
        return delta

    def _forward_with_search(self, x: torch.Tensor, ln_w: Optional[torch.Tensor], ln_b: Optional[torch.Tensor],
                             eps: float) -> Optional[torch.Tensor]:
        """Single-rank fast path: period search and block in one library call (``ftn_timesblock_forward``), so the
        period-independent first 1x1 stage runs beside the one-CTA selection kernel.  None = not applicable; the
        caller then runs the selector and the block separately (identical results)."""
        sel = self.period_selector
        if not isinstance(sel, FFTPeriodSelector) or not self._is_native_bank():
            return None
        if os.getenv("TIMES_PERIOD_MAX_UNIQ") or os.getenv("TIMES_PERIOD_BINNING"):
            return None
        B, L, C = x.shape
        if sel.k <= 0 or L <= 1 or x.dtype != torch.bfloat16:
            return None
        k = min(sel.k, L // 2)                                       # timesnet.py:122-126 (nbins - 1)
        if k <= 0 or k > nv.FTN_MAX_K:
            return None
        world = sel._world()[1]
        if world != 1 and sel.peer_comm is None:                     # NCCL transport: search and block stay separate calls
            return None
        if self.inception[0].proj.weight.device != x.device:
            self.inception = self.inception.to(x.device)
        pa = self.inception[0].packed(x.device)
        pb = self.inception[2].packed(x.device)
        # one plan buffer per (block, device, stream), zeroed once: the search rewrites it completely and leaves its
        # ticket zero, so no fill kernel per call (never a buffer born inside a graph capture)
        bufs = self.__dict__.setdefault("_plan_bufs", {})
        pkey = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
        plan_buf = bufs.get(pkey)
        if plan_buf is None and not torch.cuda.is_current_stream_capturing():
            plan_buf = bufs[pkey] = nv.new_plan(x.device)
        res = nv.timesblock_forward(x, k, sel.pmax, sel.min_period_threshold, pa.struct, pb.struct,
                                    _act_code(self._activation_name), ln_w, ln_b, eps, sel.peer_comm if world > 1 else None,
                                    plan=plan_buf)
        if res is None:
            return None
        out, plan_dev, amps, weights = res
        plan = PeriodPlan(plan_dev, amps, weights, k)
        sel._last_plan = plan
        sel._empty_device = x.device
        self._last_plan = plan
        self._last_raw = -1
        self._vec_calls += 1
        return out

    def _run(self, x: torch.Tensor, ln_w: Optional[torch.Tensor], ln_b: Optional[torch.Tensor],
             eps: float) -> torch.Tensor:
        if x.ndim != 3:
            raise ValueError("TimesBlock expects input shaped [B, L, d_model]")
        if self.period_selector is None:
            raise RuntimeError("TimesBlock.period_selector has not been set")
        x = nv.require_cuda(x, "x")
        nv.dtype_code(x.dtype)
        if _wants_grad(self, x) and self.inception is not None and self._is_native_bank():
            return self._run_differentiable(x, ln_w, ln_b, eps)
        _forward_only_guard(self, x, self._dropout)
        self._period_calls = getattr(self, "_period_calls", 0) + 1
        if self.inception is None:
            if self._configured_d_model is not None and x.size(-1) != self._configured_d_model:
                raise ValueError("Configured d_model does not match the incoming channel dimension")
            self._build_layers(x.size(-1), device=x.device, dtype=x.dtype)
        elif self.d_model is not None and x.size(-1) != self.d_model:
            raise ValueError("Number of channels changed between calls")
        B, L, C = x.shape
        with torch.no_grad():
            fused = self._forward_with_search(x, ln_w, ln_b, eps)
            if fused is not None:
                return fused
            plan = self._plan_for(x)
            if plan is None:
                if ln_w is None:
                    return x                                    # identity block (timesnet.py:797, :817)
                return nv.layer_norm(x, ln_w, ln_b, eps)
            self._vec_calls += 1
            if self._is_native_bank():
                if self.inception[0].proj.weight.device != x.device:
                    self.inception = self.inception.to(x.device)
                pa = self.inception[0].packed(x.device)
                pb = self.inception[2].packed(x.device)
                max_groups = max(1, min(plan.k, nv.FTN_MAX_K))
                nbytes = nv.inception_workspace_bytes(B, L, max_groups, pa.struct, pb.struct)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
                out = torch.empty_like(x)
                if nv.timesblock_fused(x, plan.plan_dev, max_groups, pa.struct, pb.struct,
                                       _act_code(self._activation_name), plan.weights, ln_w, ln_b, eps, out, ws):
                    return out                                  # chain + aggregation + LayerNorm, no deltas in HBM
                delta = torch.empty(max_groups, B, L, C, dtype=x.dtype, device=x.device)
                nv.period_conv(x, plan.plan_dev, max_groups, pa.struct, pb.struct, _act_code(self._activation_name),
                               delta, ws)
            else:
                delta = self._custom_bank_delta(x, plan)
            out = torch.empty_like(x)
            nv.aggregate(x, delta, plan.weights, plan.plan_dev, ln_w, ln_b, eps, out)
            return out

    def _run_differentiable(self, x: torch.Tensor, ln_w: Optional[torch.Tensor], ln_b: Optional[torch.Tensor],
                            eps: float) -> torch.Tensor:
        """Differentiable route (backward slice 2; fp32): the period search runs as usual (no gradient -- top-k has
        none), then, per period group, the fold is a padded VIEW of x (memory plumbing, torch autograd), the two
        InceptionBlocks and the activations are libflowtimes kernels forward and backward
        (``autograd.inception_block`` / ``activation``), and the weighted sum + residual is ``autograd.aggregate``.
        With the FFT selector the group weights are differentiable too (``autograd.period_weights``: softmax <- amplitude
        <- the median channel's DFT bin), the second path into x that the reference's autograd follows
        (timesnet.py:992-1009); amplitudes of a custom selector module are taken as constants."""
        from ..autograd import activation, aggregate, inception_block, layer_norm, period_weights
        if x.dtype != torch.float32:
            raise RuntimeError("TimesBlock: the differentiable route is fp32 (cast the input; bf16 stacks are forward-only)")
        if self.training and self._dropout > 0.0:
            raise RuntimeError("TimesBlock: the differentiable route has no dropout (set dropout=0 or call .eval())")
        self._period_calls = getattr(self, "_period_calls", 0) + 1
        B, L, C = x.shape
        with torch.no_grad():
            plan = self._plan_for(x.detach())
        if plan is None:
            return x if ln_w is None else layer_norm(x, ln_w, ln_b, eps)
        self._vec_calls += 1
        if self.inception[0].proj.weight.device != x.device:
            self.inception = self.inception.to(x.device)
        h = plan.host()                                          # one sync: the group geometry shapes the views below
        blk_a, blk_b = self.inception[0], self.inception[2]
        slots = max(1, min(plan.k, nv.FTN_MAX_K))
        deltas = []
        for g in range(h.n_groups):
            p, pad, cyc = int(h.grp_period[g]), int(h.grp_pad[g]), int(h.grp_cycles[g])
            grid = torch.nn.functional.pad(x, (0, 0, 0, pad)) if pad else x          # [B, L + pad, C]
            grid = grid.reshape(B, cyc, p, C).permute(0, 3, 1, 2)                     # the reference's NCHW grid
            y = inception_block(grid, blk_a)
            y = activation(y, self._activation_name)
            y = inception_block(y, blk_b)
            deltas.append((y - grid).permute(0, 2, 3, 1).reshape(B, cyc * p, C)[:, :L])
        zero = x.new_zeros(B, L, C)
        delta = torch.stack(deltas + [zero] * (slots - len(deltas)), dim=0).contiguous()
        if isinstance(self.period_selector, FFTPeriodSelector) and plan.amps is not None:
            weights = period_weights(x, plan)
        else:
            weights = plan.weights.detach()
        out = aggregate(x, delta, weights, plan.plan_dev)
        return out if ln_w is None else layer_norm(out, ln_w, ln_b, eps)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``x + sum_g w_g * (inception(fold_g(x)) - fold_g(x))``  (timesnet.py:767-818)."""
        return self._run(x, None, None, 0.0)

    def forward_norm_differentiable(self, x: torch.Tensor, norm: nn.LayerNorm) -> torch.Tensor:
        """``forward_norm`` on the differentiable route with gradients into the LayerNorm parameters as well (the fused
        forward-only route takes detached copies of them)."""
        return self._run_differentiable(x, norm.weight, norm.bias, float(norm.eps))

    def forward_norm(self, x: torch.Tensor, norm: nn.LayerNorm) -> torch.Tensor:
        """Block + inter-block residual + shared LayerNorm in one pass (timesnet.py:2058-2061)."""
        if _wants_grad(self, x) and self.inception is not None and self._is_native_bank():
            return self.forward_norm_differentiable(x, norm)
        w = norm.weight.detach().to(device=x.device, dtype=torch.float32)
        b = norm.bias.detach().to(device=x.device, dtype=torch.float32)
        return self._run(x, w.contiguous(), b.contiguous(), float(norm.eps))


# --------------------------------------------------------------------------- #
# embeddings / norms
# --------------------------------------------------------------------------- #
class PositionalEmbedding(nn.Module):
    """Sinusoidal table (timesnet.py:1104-1129); built once per (L, device)."""

    def __init__(self, d_model: int) -> None:
        super().__init__()
        self.d_model = int(d_model)

    def table(self, L: int, device: torch.device) -> torch.Tensor:
        pos = torch.arange(L, device=device, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, self.d_model, 2, device=device, dtype=torch.float32)
                        * (-math.log(10000.0) / self.d_model))
        pe = torch.zeros(L, self.d_model, device=device, dtype=torch.float32)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div[: pe[:, 1::2].shape[1]])
        return pe

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim != 3:
            raise ValueError("PositionalEmbedding expects input shaped [B, L, C]")
        B, L, _ = x.shape
        return self.table(L, x.device).to(x.dtype).unsqueeze(0).expand(B, -1, -1)


class RMSNorm(nn.Module):
    """Root-mean-square normalisation with affine parameters (timesnet.py:1132-1159); ``ftn_rms_norm``."""

    def __init__(self, d_model: int, eps: float = 1e-5) -> None:
        super().__init__()
        if d_model <= 0:
            raise ValueError("RMSNorm expects a positive embedding dimension")
        self.eps = float(eps)
        self.weight = nn.Parameter(torch.ones(int(d_model)))
        self.bias = nn.Parameter(torch.zeros(int(d_model)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.size(-1) != self.weight.numel():
            raise ValueError("RMSNorm dimension mismatch")
        x = nv.require_cuda(x, "x")
        nv.dtype_code(x.dtype)
        with torch.no_grad():
            return nv.rms_norm(x, self.weight.detach().to(device=x.device, dtype=torch.float32).contiguous(),
                               self.bias.detach().to(device=x.device, dtype=torch.float32).contiguous(), self.eps)


class DataEmbedding(nn.Module):
    """Value + positional (+ temporal) embedding (timesnet.py:1200-1325).

    value GEMM and the ``value + gate * LN(aux)`` combine run in libflowtimes;
    without time marks ``LN(PE)`` is batch independent and cached per length.
    ``out_dtype`` lets TimesNet emit bf16 features for the bf16 stack directly.
    """

    _VALID_NORM_MODES = {"none", "layer", "rms", "decoupled"}

    def __init__(self, c_in: int, d_model: int, dropout: float, time_features: Optional[int] = None,
                 use_norm: bool = True, embed_norm_mode: Optional[str] = None) -> None:
        super().__init__()
        self.value_embedding = nn.Linear(int(c_in), int(d_model))
        self.position_embedding = PositionalEmbedding(d_model)
        self.temporal_embedding = (nn.Linear(int(time_features), int(d_model))
                                   if time_features is not None and time_features > 0 else None)
        if embed_norm_mode is None:
            embed_norm_mode = "decoupled" if bool(use_norm) else "none"
        mode = embed_norm_mode.lower()
        if mode not in self._VALID_NORM_MODES:
            raise ValueError(
                f"embed_norm_mode must be one of {sorted(self._VALID_NORM_MODES)}, got {embed_norm_mode!r}")
        self.embed_norm_mode = mode
        self.use_norm = mode != "none"
        self.norm: Optional[nn.Module] = None
        self.aux_norm: Optional[nn.Module] = None
        if mode == "layer":
            self.norm = nn.LayerNorm(int(d_model))
            self.register_parameter("gate", None)
        elif mode == "rms":
            self.norm = RMSNorm(int(d_model))
            self.register_parameter("gate", None)
        elif mode == "decoupled":
            self.aux_norm = nn.LayerNorm(int(d_model))
            self.gate = nn.Parameter(torch.full((1, 1, int(d_model)), 0.1, dtype=torch.float32))
        else:
            self.register_parameter("gate", None)
        self.dropout = nn.Dropout(float(dropout))
        self._aux_cache = None
        self._ve_cache = None

    def _value_weight_split(self, device) -> torch.Tensor:
        """value_embedding.weight as the three-plane bf16 operand of ``ftn_embed_tc`` (re-packed when it changes)."""
        w = self.value_embedding.weight
        key = (w.data_ptr(), w._version, str(device))
        if self._ve_cache is None or self._ve_cache[0] != key:
            kp = (w.shape[1] + 15) // 16 * 16
            self._ve_cache = (key, split_linear_weight(w.detach().to(device=device, dtype=torch.float32), kp))
        return self._ve_cache[1]

    def _aux_static(self, L: int, device: torch.device) -> torch.Tensor:
        norm_key = (None if self.aux_norm is None
                    else (self.aux_norm.weight.data_ptr(), self.aux_norm.weight._version,
                          self.aux_norm.bias._version))
        key = (L, str(device), norm_key)
        if self._aux_cache is None or self._aux_cache[0] != key:
            pe = self.position_embedding.table(L, device)
            if self.embed_norm_mode == "decoupled":
                pe = nv.layer_norm(pe, self.aux_norm.weight.detach().float().contiguous(),
                                   self.aux_norm.bias.detach().float().contiguous(), self.aux_norm.eps)
            self._aux_cache = (key, pe.contiguous())
        return self._aux_cache[1]

    def forward(self, x: torch.Tensor, x_mark: Optional[torch.Tensor] = None,
                out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        if x.ndim not in (3, 4):
            raise ValueError("DataEmbedding expects input shaped [B, L, C] or [B, L, N, C]")
        shape4 = None
        if x.ndim == 4:
            B4, L4, N4, C4 = x.shape
            shape4 = (B4, L4, N4)
            x = x.reshape(B4 * N4, L4, C4)
            if x_mark is not None:
                if x_mark.ndim == 3:
                    if x_mark.shape[0] != B4 or x_mark.shape[1] != L4:
                        raise ValueError("x_mark must match batch/time dimensions of x")
                    x_mark = x_mark.unsqueeze(2).expand(-1, -1, N4, -1)
                elif x_mark.ndim != 4 or tuple(x_mark.shape[:3]) != (B4, L4, N4):
                    raise ValueError("x_mark must have shape [B, L, T] or [B, L, N, T]")
                x_mark = x_mark.reshape(B4 * N4, L4, x_mark.size(-1))
        elif x_mark is not None and x_mark.ndim != 3:
            raise ValueError("x_mark must share dimensions [B, L, T]")
        x = nv.require_cuda(x, "x").to(torch.float32)
        B, L, _ = x.shape
        dev = x.device
        out_dtype = out_dtype or x.dtype
        with torch.no_grad():
            ve = self.value_embedding
            d_model = ve.out_features
            batched = self.temporal_embedding is not None and x_mark is not None
            if batched:
                te = self.temporal_embedding
                mark = nv.require_cuda(x_mark, "x_mark").to(torch.float32)
                aux = nv.linear(mark, te.weight.detach().float().contiguous(), te.bias.detach().float().contiguous())
                aux = aux + self.position_embedding.table(L, dev).unsqueeze(0)
                if self.embed_norm_mode == "decoupled":
                    aux = nv.layer_norm(aux.contiguous(), self.aux_norm.weight.detach().float().contiguous(),
                                        self.aux_norm.bias.detach().float().contiguous(), self.aux_norm.eps)
                aux = aux.contiguous()
            else:
                aux = self._aux_static(L, dev)
            if self.embed_norm_mode == "decoupled":
                gate = self.gate.detach().float().reshape(-1).contiguous()
            else:
                gate = torch.ones(d_model, dtype=torch.float32, device=dev)
            fused_dtype = torch.float32 if self.embed_norm_mode in ("layer", "rms") else out_dtype
            # K0: value GEMM on the tensor cores with the combine fused into its epilogue (None = shape not eligible)
            out = nv.embed_tc(x.contiguous(), self._value_weight_split(dev), ve.bias.detach().float().contiguous(), aux,
                              batched, gate, fused_dtype)
            if out is not None:
                if self.embed_norm_mode in ("layer", "rms"):
                    norm_fn = nv.layer_norm if self.embed_norm_mode == "layer" else nv.rms_norm
                    out = norm_fn(out, self.norm.weight.detach().float().contiguous(),
                                  self.norm.bias.detach().float().contiguous(), self.norm.eps).to(out_dtype)
                if shape4 is not None:
                    return out.view(shape4[0], shape4[1], shape4[2], out.size(-1))
                return out
            value = nv.linear(x, ve.weight.detach().float().contiguous(), ve.bias.detach().float().contiguous())
            if self.embed_norm_mode in ("layer", "rms"):                       # timesnet.py:1313-1316
                out = nv.embed_combine(value, aux, gate, batched, torch.float32)
                norm_fn = nv.layer_norm if self.embed_norm_mode == "layer" else nv.rms_norm
                out = norm_fn(out, self.norm.weight.detach().float().contiguous(),
                              self.norm.bias.detach().float().contiguous(), self.norm.eps).to(out_dtype)
            else:
                out = nv.embed_combine(value, aux, gate, batched, out_dtype)
        if shape4 is not None:
            return out.view(shape4[0], shape4[1], shape4[2], out.size(-1))      # timesnet.py:1323-1324
        return out


class LowRankTemporalContext(nn.Module):
    """Zero-mean low-rank temporal context (timesnet.py:1328-1371)."""

    def __init__(self, rank: int, init_scale: float = 1e-2) -> None:
        super().__init__()
        if rank <= 0:
            raise ValueError("LowRankTemporalContext requires a positive rank")
        self.rank = int(rank)
        self.scale = nn.Parameter(torch.as_tensor(float(init_scale), dtype=torch.float32))
        self.register_buffer("_cached_basis", torch.empty(0), persistent=False)
        self._cached_length: int = 0

    def _compute_basis(self, length: int, device: torch.device, dtype: torch.dtype) -> torch.Tensor:
        """DCT-like basis, column-centred and L2-normalised; one-off per length (timesnet.py:1340-1351)."""
        steps = torch.arange(length, device=device, dtype=torch.float32).unsqueeze(1)
        freqs = torch.arange(1, self.rank + 1, device=device, dtype=torch.float32).unsqueeze(0)
        basis = torch.cos(math.pi / float(length) * (steps + 0.5) * freqs)
        basis = basis - basis.mean(dim=0, keepdim=True)
        norm = torch.linalg.norm(basis, dim=0, keepdim=True)
        return (basis / norm.clamp_min(torch.finfo(torch.float32).eps)).to(dtype)

    def _basis(self, length: int, reference: torch.Tensor) -> torch.Tensor:
        if (self._cached_basis.numel() == 0 or self._cached_length != length
                or self._cached_basis.device != reference.device):
            self._cached_basis = self._compute_basis(length, reference.device, torch.float32).detach().contiguous()
            self._cached_length = length
        return self._cached_basis

    def add_to(self, x: torch.Tensor, coeff: torch.Tensor) -> torch.Tensor:
        """``x + forward(coeff, L)`` fused in one kernel (einsum + centring + scale + add)."""
        if coeff.ndim != 3:
            raise ValueError("LowRankTemporalContext expects coeff shaped [B, N, R]")
        if coeff.size(-1) != self.rank:
            raise ValueError("Coefficient dimension mismatch with configured rank")
        x = nv.require_cuda(x, "x").to(torch.float32)
        coeff = nv.require_cuda(coeff, "coeff").to(torch.float32)
        out = torch.empty_like(x)
        with torch.no_grad():
            nv.context_add(x, coeff, self._basis(x.shape[1], x),
                           self.scale.detach().to(device=x.device, dtype=torch.float32).reshape(1), out)
        return out

    def forward(self, coeff: torch.Tensor, length: int) -> torch.Tensor:
        if coeff.ndim != 3:
            raise ValueError("LowRankTemporalContext expects coeff shaped [B, N, R]")
        if coeff.size(-1) != self.rank:
            raise ValueError("Coefficient dimension mismatch with configured rank")
        coeff = nv.require_cuda(coeff, "coeff")
        zeros = torch.zeros(coeff.shape[0], int(length), coeff.shape[1], dtype=torch.float32, device=coeff.device)
        return self.add_to(zeros, coeff).to(coeff.dtype)


# --------------------------------------------------------------------------- #
# TimesNet
# --------------------------------------------------------------------------- #
class TimesNet(nn.Module):
    """TimesNet forecaster, ``[B,T,N] -> (rate[B,H,N], dispersion[B,H,N])`` (timesnet.py:1374-2102).

    Extra keyword (not in the reference): ``stack_dtype`` -- activation dtype of
    the TimesBlock stack (``torch.bfloat16`` reproduces BASELINE.json's bf16
    configs: features are emitted in bf16 by the embedding, blocks and the shared
    LayerNorm run on bf16 activations with fp32 math, the head reads bf16).
    """

    def __init__(self, input_len: int, pred_len: int, d_model: int, n_layers: int, k_periods: int,
                 kernel_set, dropout: float, activation: str, mode: str, d_ff: Optional[int] = None,
                 bottleneck_ratio: float = 1.0, min_period_threshold: int = 1, channels_last: bool = False,
                 use_checkpoint: bool = True, use_embedding_norm: bool = True,
                 embed_norm_mode: Optional[str] = None, min_sigma: float = 1e-3, min_sigma_vector=None,
                 id_embed_dim: int = 32, static_proj_dim: Optional[int] = None, static_layernorm: bool = True,
                 use_zero_mean_context: bool = False, context_rank: int = 0, context_scale: float = 1e-2,
                 use_constant_context_bias: bool = False, use_late_bias_head: bool = True,
                 stack_dtype: Optional[torch.dtype] = None) -> None:
        super().__init__()
        del channels_last                       # kept for signature compatibility
        assert mode in ("direct", "recursive")
        self.mode = mode
        self.input_len = int(input_len)
        self.pred_len = int(pred_len)
        self.requested_d_model = int(d_model)
        if d_ff is None:
            self.requested_d_ff: Optional[int] = None
        else:
            if int(d_ff) <= 0:
                raise ValueError("d_ff must be a positive integer")
            self.requested_d_ff = int(d_ff)
        self.d_model: Optional[int] = None
        self.d_ff: Optional[int] = self.requested_d_ff
        self.bottleneck_ratio = float(bottleneck_ratio)
        if self.bottleneck_ratio <= 0:
            raise ValueError("bottleneck_ratio must be a positive value")
        self.n_layers = int(n_layers)
        self.dropout = float(dropout)
        self.use_checkpoint = bool(use_checkpoint)      # forward-only build: nothing to checkpoint
        self.use_embedding_norm = bool(use_embedding_norm)
        self.embed_norm_mode = embed_norm_mode if embed_norm_mode is not None else (
            "decoupled" if self.use_embedding_norm else "none")
        self.min_sigma = float(min_sigma)
        self.k_periods = int(k_periods)
        self.kernel_set = list(kernel_set)
        self.stack_dtype = stack_dtype
        self.check_finite = True                        # reference raises on non-finite outputs (one sync)
        self.period_selector = FFTPeriodSelector(self.k_periods, self.input_len, min_period_threshold)
        self.blocks = nn.ModuleList([
            TimesBlock(d_model=None, d_ff=self.requested_d_ff, kernel_set=self.kernel_set, dropout=self.dropout,
                       activation=activation, bottleneck_ratio=self.bottleneck_ratio)
            for _ in range(self.n_layers)])
        for idx, block in enumerate(self.blocks):
            block.block_index = idx
            object.__setattr__(block, "period_selector", self.period_selector)
        self.residual_dropout = nn.Dropout(self.dropout)
        self.layer_norm: Optional[nn.LayerNorm] = None
        self.forecast_time_proj = nn.Linear(self.input_len, self.pred_len)
        with torch.no_grad():                           # "copy the last step" prior (timesnet.py:1461-1466)
            self.forecast_time_proj.weight.zero_()
            if self.pred_len > 0:
                self.forecast_time_proj.weight[:, -1] = 1.0
            self.forecast_time_proj.bias.zero_()
        self.embedding: Optional[DataEmbedding] = None
        self.embedding_time_features: Optional[int] = None
        self.mu_head: Optional[nn.Linear] = None
        self.sigma_head: Optional[nn.Linear] = None
        self.output_dim: Optional[int] = None
        self.input_channels: Optional[int] = None
        self._out_steps = self.pred_len if self.mode == "direct" else 1
        self.register_buffer("min_sigma_vector", None)
        if min_sigma_vector is not None:
            self.min_sigma_vector = torch.as_tensor(min_sigma_vector, dtype=torch.float32).reshape(1, 1, -1)
        self.id_embed_dim = int(id_embed_dim)
        if self.id_embed_dim < 0:
            raise ValueError("id_embed_dim must be non-negative")
        if static_proj_dim is None:
            self.static_proj_dim: Optional[int] = None
        else:
            if int(static_proj_dim) <= 0:
                raise ValueError("static_proj_dim must be a positive integer when provided")
            self.static_proj_dim = int(static_proj_dim)
        self.static_layernorm = bool(static_layernorm)
        self.series_embedding: Optional[nn.Embedding] = None
        self.static_proj: Optional[nn.Linear] = None
        self.static_norm: Optional[nn.Module] = None
        self.context_norm: Optional[nn.LayerNorm] = None
        self.context_proj: Optional[nn.Linear] = None
        self.context_coeff: Optional[nn.Linear] = None
        self.temporal_context: Optional[LowRankTemporalContext] = None
        self.late_bias_norm: Optional[nn.LayerNorm] = None
        self.late_bias_head: Optional[nn.Linear] = None
        self.register_parameter("late_bias_gate", None)
        self.pre_embedding_norm: Optional[nn.Module] = None
        self.pre_embedding_dropout = nn.Dropout(self.dropout)
        self._static_in_features: Optional[int] = None
        self._static_out_dim: int = 0
        self._series_id_vocab: Optional[int] = None
        self._series_id_reference: Optional[torch.Tensor] = None
        self.debug_memory: bool = False
        self.use_zero_mean_context = bool(use_zero_mean_context)
        self.use_constant_context_bias = bool(use_constant_context_bias)
        self.use_late_bias_head = bool(use_late_bias_head)
        self.context_rank = int(context_rank)
        if self.context_rank < 0:
            raise ValueError("context_rank must be non-negative")
        self.context_scale_default = float(context_scale)

    # ------------------------------------------------------------------ #
    @staticmethod
    def _zero_linear(i: int, o: int, dev) -> nn.Linear:
        lin = nn.Linear(i, o)
        with torch.no_grad():
            lin.weight.zero_()
            lin.bias.zero_()
        return lin.to(dev)

    def _ensure_embedding(self, x: torch.Tensor, x_mark=None, series_static=None, series_ids=None) -> None:
        """Lazy construction of every N-dependent sub-module (timesnet.py:1514-1849)."""
        dev = x.device
        c_in = int(x.size(-1))
        time_dim = int(x_mark.size(-1)) if x_mark is not None else 0
        if self.input_channels is None:
            self.input_channels = c_in
        elif self.input_channels != c_in:
            raise ValueError("Number of series changed between calls")
        if self.d_model is None:
            self.d_model = int(self.requested_d_model)
        self.d_ff = self.d_model if self.requested_d_ff is None else self.requested_d_ff

        static_out = 0
        if series_static is not None:
            if series_static.ndim == 2:
                ref = series_static
            elif series_static.ndim == 3:
                ref = series_static[0]
            else:
                raise ValueError("series_static must have shape [N, F] or [B, N, F]")
            if ref.size(0) != c_in:
                raise ValueError("series_static must align with the number of input series")
            static_in = int(ref.size(-1))
            if static_in <= 0:
                raise ValueError("series_static must have at least one feature")
            if self.static_proj is None:
                proj_dim = self.static_proj_dim if self.static_proj_dim is not None else static_in
                self.static_proj = nn.Linear(static_in, proj_dim).to(dev)
                self.static_norm = nn.LayerNorm(proj_dim).to(dev) if self.static_layernorm else nn.Identity()
                self._static_in_features = static_in
            elif self.static_proj.in_features != static_in:
                raise ValueError("series_static feature dimension changed between calls")
            static_out = int(self.static_proj.out_features)
        elif self.static_proj is not None:
            static_out = int(self.static_proj.out_features)
        self._static_out_dim = static_out

        id_dim = 0
        if self.id_embed_dim > 0:
            ids_ref = None
            if series_ids is not None:
                if series_ids.ndim == 1:
                    ids_ref = series_ids.to(torch.long)
                elif series_ids.ndim == 2:
                    ids_ref = series_ids[0].to(torch.long)
                else:
                    raise ValueError("series_ids must have shape [N] or [B, N]")
                if ids_ref.numel() != c_in:
                    raise ValueError("series_ids length must match number of series")
            if self.series_embedding is None:
                if ids_ref is None:
                    ids_ref = torch.arange(c_in, device=dev, dtype=torch.long)
                vocab = int(ids_ref.max().item()) + 1 if ids_ref.numel() > 0 else c_in
                self.series_embedding = nn.Embedding(vocab, self.id_embed_dim).to(dev)
                self._series_id_vocab = vocab
                self._series_id_reference = ids_ref.to(dev)
            else:
                if ids_ref is not None:
                    # the vocabulary check reads the ids back (one host sync): done once per distinct ids tensor, so a
                    # caller that passes the same tensor every step stays sync-free (and capturable in a CUDA graph)
                    key = (series_ids.data_ptr(), series_ids._version, tuple(series_ids.shape))
                    if getattr(self, "_ids_checked", None) != key:
                        vocab = int(ids_ref.max().item()) + 1 if ids_ref.numel() > 0 else c_in
                        if vocab > int(self.series_embedding.num_embeddings):
                            raise ValueError("series_ids vocabulary expanded between calls")
                        self._ids_checked = key
                    self._series_id_reference = ids_ref.to(dev)
                elif self._series_id_reference is None:
                    self._series_id_reference = torch.arange(c_in, device=dev, dtype=torch.long)
                self._series_id_vocab = int(self.series_embedding.num_embeddings)
            if self._series_id_reference is not None and self._series_id_reference.numel() != c_in:
                raise ValueError("series identifier count changed between calls")
            id_dim = int(self.series_embedding.embedding_dim)

        ctx_dim = static_out + id_dim
        if ctx_dim > 0:
            if self.context_norm is None or tuple(self.context_norm.normalized_shape) != (ctx_dim,):
                self.context_norm = nn.LayerNorm(ctx_dim).to(dev)
            if self.use_zero_mean_context and self.context_rank > 0:
                if (self.context_coeff is None or self.context_coeff.in_features != ctx_dim
                        or self.context_coeff.out_features != self.context_rank):
                    self.context_coeff = self._zero_linear(ctx_dim, self.context_rank, dev)
                if self.temporal_context is None or self.temporal_context.rank != self.context_rank:
                    self.temporal_context = LowRankTemporalContext(self.context_rank,
                                                                   self.context_scale_default).to(dev)
            else:
                self.context_coeff = None
                self.temporal_context = None
            if self.use_constant_context_bias:
                if self.context_proj is None or self.context_proj.in_features != ctx_dim:
                    self.context_proj = self._zero_linear(ctx_dim, 1, dev)
            else:
                self.context_proj = None
            if self.use_late_bias_head:
                if self.late_bias_norm is None or tuple(self.late_bias_norm.normalized_shape) != (ctx_dim,):
                    self.late_bias_norm = nn.LayerNorm(ctx_dim).to(dev)
                if (self.late_bias_head is None or self.late_bias_head.in_features != ctx_dim
                        or self.late_bias_head.out_features != self._out_steps):
                    self.late_bias_head = self._zero_linear(ctx_dim, self._out_steps, dev)
                shape = (1, self._out_steps, 1)
                if not isinstance(self.late_bias_gate, nn.Parameter) or tuple(self.late_bias_gate.shape) != shape:
                    self.late_bias_gate = nn.Parameter(torch.full(shape, 0.05, dtype=torch.float32, device=dev))
            else:
                self.late_bias_norm = None
                self.late_bias_head = None
                if isinstance(self.late_bias_gate, nn.Parameter):
                    self.late_bias_gate = None
        else:
            self.context_norm = self.context_proj = self.context_coeff = None
            self.temporal_context = self.late_bias_norm = self.late_bias_head = None
            if isinstance(self.late_bias_gate, nn.Parameter):
                self.late_bias_gate = None

        # allocated by the reference but never used in forward; kept so checkpoints load
        if 1 + ctx_dim <= 1:
            if not isinstance(self.pre_embedding_norm, nn.Identity):
                self.pre_embedding_norm = nn.Identity()
        elif (not isinstance(self.pre_embedding_norm, nn.LayerNorm)
              or tuple(self.pre_embedding_norm.normalized_shape) != (1 + ctx_dim,)):
            self.pre_embedding_norm = nn.LayerNorm(1 + ctx_dim).to(dev)

        if isinstance(self.min_sigma_vector, torch.Tensor) and self.min_sigma_vector.numel() > 0:
            cur = int(self.min_sigma_vector.shape[-1])
            if cur < c_in:
                raise ValueError("min_sigma_vector length does not match number of series")
            if cur != c_in:
                self.min_sigma_vector = self.min_sigma_vector[..., :c_in]
        if self.embedding_time_features is not None and self.embedding_time_features != time_dim:
            raise ValueError("Temporal feature dimension changed between calls")
        if self.embedding is None:
            self.embedding = DataEmbedding(c_in, self.d_model, self.dropout,
                                           time_features=time_dim if time_dim > 0 else None,
                                           use_norm=self.use_embedding_norm,
                                           embed_norm_mode=self.embed_norm_mode).to(dev)
        self.embedding_time_features = time_dim
        if self.layer_norm is None or tuple(self.layer_norm.normalized_shape) != (self.d_model,):
            self.layer_norm = nn.LayerNorm(self.d_model).to(dev)
        if self.mu_head is None or self.mu_head.in_features != self.d_model or self.mu_head.out_features != c_in:
            self.mu_head = self._zero_linear(self.d_model, c_in, dev)        # zero-init (timesnet.py:1824-1826)
        if (self.sigma_head is None or self.sigma_head.in_features != self.d_model
                or self.sigma_head.out_features != c_in):
            self.sigma_head = self._zero_linear(self.d_model, c_in, dev)     # zero-init (timesnet.py:1840-1842)
        self.output_dim = self.input_channels
        if next(self.parameters()).device != dev or self.forecast_time_proj.weight.device != dev:
            self.to(dev)

    # ------------------------------------------------------------------ #
    @staticmethod
    def _f32(t: torch.Tensor) -> torch.Tensor:
        return t.detach().to(torch.float32).contiguous()

    def _heads_split(self, dev):
        """[mu_head; sigma_head] as ONE three-plane bf16 operand of ``ftn_nb_head_tc`` (re-packed when a head changes):
        rows [0, N) = mu_head.weight, rows [Np, Np + N) = sigma_head.weight, Np = N rounded up to 128."""
        ps = (self.mu_head.weight, self.mu_head.bias, self.sigma_head.weight, self.sigma_head.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps) + (str(dev),)
        cache = getattr(self, "_heads_cache", None)
        if cache is None or cache[0] != key:
            N, Cc = self.mu_head.weight.shape
            n_pad = (N + 127) // 128 * 128
            f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32)
            w = split_linear_weight(f32(self.mu_head.weight), Cc, 2 * n_pad, 0)
            w += split_linear_weight(f32(self.sigma_head.weight), Cc, 2 * n_pad, n_pad)
            b = torch.zeros(2 * n_pad, dtype=torch.float32, device=dev)
            b[:N] = f32(self.mu_head.bias)
            b[n_pad:n_pad + N] = f32(self.sigma_head.bias)
            self._heads_cache = (key, w.contiguous(), b, n_pad)
        return self._heads_cache[1], self._heads_cache[2], self._heads_cache[3]

    def _time_proj_split(self, Wt: torch.Tensor, steps: int, dev) -> Optional[torch.Tensor]:
        """The last ``steps`` rows of forecast_time_proj.weight as the three-plane bf16 operand of the tensor-core time
        projection (re-packed when the weight changes).  None while a CUDA graph is being captured before the first pack
        (the SIMT projection is captured instead) or when the stack is fp32."""
        if self.stack_dtype != torch.bfloat16:
            return None
        p = self.forecast_time_proj.weight
        key = (p.data_ptr(), p._version, steps, str(dev))
        cache = getattr(self, "_time_proj_cache", None)
        if cache is None or cache[0] != key:
            if torch.cuda.is_current_stream_capturing():
                return None
            packed = nv.time_proj_pack(Wt)
            torch.cuda.current_stream(dev).synchronize()
            self._time_proj_cache = (key, packed)
        return self._time_proj_cache[1]

    def _context_key(self, B: int, N: int, dev, series_static, series_ids):
        """Identity of everything the per-series context and the late bias depend on (they do NOT depend on x): parameter
        versions and the id tensor.  None = not cacheable (static features are per call)."""
        if series_static is not None:
            return None
        mods = (self.series_embedding, self.context_norm, self.late_bias_norm, self.late_bias_head)
        ps = [p for m in mods if isinstance(m, nn.Module) for p in m.parameters()]
        # the id tensor is matched by OBJECT identity + version (the cache keeps it alive, so its address cannot be
        # recycled for a different tensor); a fresh tensor per call simply misses
        ids = None if series_ids is None else (id(series_ids), series_ids._version)
        ref = self._series_id_reference
        ref_key = None if ref is None else (ref.data_ptr(), ref._version)
        return (B, N, str(dev), ids, ref_key if series_ids is None else None,
                tuple((p.data_ptr(), p._version) for p in ps))

    def _context(self, B: int, N: int, dev, series_static, series_ids) -> Optional[torch.Tensor]:
        """Static projection + id embedding + context LayerNorm (timesnet.py:1886-1957)."""
        key = self._context_key(B, N, dev, series_static, series_ids)
        cache = getattr(self, "_ctx_cache", None)
        if key is not None and cache is not None and cache[0] == key:
            return cache[1]
        ctx = self._context_compute(B, N, dev, series_static, series_ids)
        # re-key after the call (it may have stored the id reference); never cache tensors born inside a graph capture
        if key is not None and ctx is not None and not torch.cuda.is_current_stream_capturing():
            self._ctx_cache = (self._context_key(B, N, dev, series_static, series_ids), ctx, series_ids)
            self._late_cache = None
        return ctx

    def _late_bias(self, ctx: torch.Tensor) -> torch.Tensor:
        """late_bias_head(late_bias_norm(ctx)) (timesnet.py:2028-2040); cached with the context it was computed from."""
        cache = getattr(self, "_late_cache", None)
        ctx_cache = getattr(self, "_ctx_cache", None)
        cached_ctx = ctx_cache is not None and ctx_cache[1] is ctx
        if cached_ctx and cache is not None and cache[0] is ctx:
            return cache[1]
        c = nv.layer_norm(ctx, self._f32(self.late_bias_norm.weight), self._f32(self.late_bias_norm.bias),
                          self.late_bias_norm.eps)
        late = nv.linear(c, self._f32(self.late_bias_head.weight), self._f32(self.late_bias_head.bias))
        if cached_ctx and not torch.cuda.is_current_stream_capturing():
            self._late_cache = (ctx, late)
        return late

    def _late_bias_gated_t(self, late: torch.Tensor, gate: torch.Tensor) -> Optional[torch.Tensor]:
        """``gate[h] * late[b, n, h]`` as a step-major ``[B, steps, N]`` tensor (the layout of the head's output, so the
        tensor-core head reads it coalesced) -- a cached repack of the cached late bias, like the weight packs.  None
        when the late bias is not the cached one (per-call static features)."""
        cache = getattr(self, "_late_cache", None)
        if cache is None or cache[1] is not late or torch.cuda.is_current_stream_capturing() and len(cache) < 4:
            return None
        gkey = (self.late_bias_gate.data_ptr(), self.late_bias_gate._version)
        if len(cache) < 4 or cache[2] != gkey:
            packed = (late * gate.view(1, 1, -1)).transpose(1, 2).contiguous()
            self._late_cache = (cache[0], late, gkey, packed)
        return self._late_cache[3]

    def _context_compute(self, B: int, N: int, dev, series_static, series_ids) -> Optional[torch.Tensor]:
        comps = []
        if self.static_proj is not None and series_static is not None:
            if series_static.ndim == 2:
                st = series_static.unsqueeze(0).expand(B, -1, -1)
            elif series_static.ndim == 3:
                if series_static.size(0) != B:
                    raise ValueError("series_static batch dimension must match input batch size")
                st = series_static
            else:
                raise ValueError("series_static must have shape [N, F] or [B, N, F]")
            st = st.to(device=dev, dtype=torch.float32).contiguous()
            sp = nv.linear(st, self._f32(self.static_proj.weight), self._f32(self.static_proj.bias))
            if isinstance(self.static_norm, nn.LayerNorm):
                sp = nv.layer_norm(sp, self._f32(self.static_norm.weight), self._f32(self.static_norm.bias),
                                   self.static_norm.eps)
            comps.append(sp)
        if self.series_embedding is not None and self.id_embed_dim > 0:
            if series_ids is None:
                if self._series_id_reference is None:
                    ids = torch.arange(N, device=dev, dtype=torch.long).unsqueeze(0)
                else:
                    ids = self._series_id_reference.view(1, -1).to(dev)
                    if ids.size(1) != N:
                        raise ValueError("Stored series identifiers do not match input dimension")
            else:
                ids = series_ids
                if ids.ndim == 1:
                    ids = ids.unsqueeze(0)
                if ids.ndim != 2:
                    raise ValueError("series_ids must have shape [N] or [B, N]")
                if ids.size(0) not in (1, B):
                    raise ValueError("series_ids batch dimension does not match input")
                if ids.size(1) != N:
                    raise ValueError("series_ids length must match number of series")
                ids = ids.to(device=dev, dtype=torch.long)
                self._series_id_reference = ids[0].detach().clone()
            if ids.size(0) == 1 and B > 1:
                ids = ids.expand(B, -1)
            comps.append(self.series_embedding.weight.detach()[ids])       # row gather (memory plumbing)
        if not comps:
            return None
        ctx = torch.cat(comps, dim=-1).contiguous()
        if self.context_norm is not None:
            ctx = nv.layer_norm(ctx, self._f32(self.context_norm.weight), self._f32(self.context_norm.bias),
                                self.context_norm.eps)
        return ctx

    def _forward_differentiable(self, x, x_mark, series_static, series_ids):
        """Training route (``model.differentiable = True`` or an input that requires grad; fp32 stack, dropout 0): the
        same forward assembled from the differentiable building blocks of ``timesnet_forecast.autograd`` -- every dense
        layer, LayerNorm, convolution, activation, the aggregation, the period weights and the NB head run in
        libflowtimes kernels forward AND backward; torch autograd only chains them (views, ``cat``, the id gather and a
        few broadcast adds / multiplies are its own ops).  The period search itself has no gradient (top-k)."""
        from .. import autograd as ag
        if (self.stack_dtype or torch.float32) != torch.float32:
            raise RuntimeError("TimesNet: the differentiable route needs an fp32 stack (stack_dtype=None)")
        if self.training and float(self.dropout) > 0.0:
            raise RuntimeError("TimesNet: the differentiable route has no dropout (build the model with dropout=0)")
        B, T, N = x.shape
        L, dev = self.input_len, x.device
        xv = x[:, -L:, :].contiguous()
        mark = x_mark[:, -L:, :].contiguous() if x_mark is not None else None
        self._ensure_embedding(xv, mark, series_static, series_ids)
        steps = self.pred_len if self.mode == "direct" else self._out_steps
        # ---- per-series context (timesnet.py:1886-1957) ----
        comps = []
        if self.static_proj is not None and series_static is not None:
            st = series_static.unsqueeze(0).expand(B, -1, -1) if series_static.ndim == 2 else series_static
            sp = ag.linear(st.to(device=dev, dtype=torch.float32).contiguous(), self.static_proj.weight, self.static_proj.bias)
            if isinstance(self.static_norm, nn.LayerNorm):
                sp = ag.layer_norm(sp, self.static_norm.weight, self.static_norm.bias, self.static_norm.eps)
            comps.append(sp)
        if self.series_embedding is not None and self.id_embed_dim > 0:
            if series_ids is None:
                ids = (torch.arange(N, device=dev) if self._series_id_reference is None
                       else self._series_id_reference.to(dev)).view(1, -1)
            else:
                ids = series_ids.to(device=dev, dtype=torch.long)
                ids = ids.unsqueeze(0) if ids.ndim == 1 else ids
            ids = ids.expand(B, -1) if ids.size(0) == 1 and B > 1 else ids
            comps.append(self.series_embedding.weight[ids])                  # row gather: torch indexing (and its index_add)
        ctx = None
        if comps:
            ctx = torch.cat(comps, dim=-1).contiguous()
            if self.context_norm is not None:
                ctx = ag.layer_norm(ctx, self.context_norm.weight, self.context_norm.bias, self.context_norm.eps)
        feat_in = xv
        if (ctx is not None and self.use_zero_mean_context and self.context_coeff is not None
                and self.temporal_context is not None):                      # timesnet.py:1966-1983
            coeff = ag.linear(ctx, self.context_coeff.weight, self.context_coeff.bias)            # [B, N, R]
            basis = self.temporal_context._basis(L, xv)                                            # [L, R] constant
            tc = ag.linear(coeff, basis, None).permute(0, 2, 1)                                    # einsum('lr,bnr->bln')
            tc = tc - tc.mean(dim=1, keepdim=True)
            feat_in = feat_in + tc * self.temporal_context.scale.to(torch.float32)
        if ctx is not None and self.use_constant_context_bias and self.context_proj is not None:
            cb = ag.linear(ctx, self.context_proj.weight, self.context_proj.bias)
            feat_in = feat_in + cb.squeeze(-1).unsqueeze(1)
        # ---- K0: DataEmbedding (timesnet.py:1295-1320) ----
        emb = self.embedding
        value = ag.linear(feat_in.contiguous(), emb.value_embedding.weight, emb.value_embedding.bias)
        aux = emb.position_embedding.table(L, dev).unsqueeze(0)
        if emb.temporal_embedding is not None and mark is not None:
            aux = aux + ag.linear(mark.to(torch.float32), emb.temporal_embedding.weight, emb.temporal_embedding.bias)
        if emb.embed_norm_mode == "decoupled":
            auxn = ag.layer_norm(aux.expand(B, -1, -1).contiguous() if aux.size(0) == 1 else aux, emb.aux_norm.weight,
                                 emb.aux_norm.bias, emb.aux_norm.eps)
            seq = value + emb.gate.to(torch.float32) * auxn
        elif emb.embed_norm_mode == "layer":
            seq = ag.layer_norm(value + aux, emb.norm.weight, emb.norm.bias, emb.norm.eps)
        elif emb.embed_norm_mode == "none":
            seq = value + aux
        else:
            raise NotImplementedError(f"differentiable route: embed_norm_mode={emb.embed_norm_mode!r}")
        # ---- TimesBlock stack + shared LayerNorm (timesnet.py:2050-2061) ----
        for block in self.blocks:
            object.__setattr__(block, "period_selector", self.period_selector)
            seq = block.forward_norm_differentiable(seq.contiguous(), self.layer_norm)
        # ---- head (timesnet.py:2008-2014, 2063-2093) ----
        hist_steps = min(steps, L)
        hist = xv[:, -hist_steps:, :]
        if hist_steps < steps:
            hist = torch.cat([hist, hist[:, -1:, :].expand(-1, steps - hist_steps, -1)], dim=1)
        Wt = self.forecast_time_proj.weight[-steps:, :] if steps != self.pred_len else self.forecast_time_proj.weight
        bt = self.forecast_time_proj.bias[-steps:] if steps != self.pred_len else self.forecast_time_proj.bias
        late = gate = None
        if (ctx is not None and self.late_bias_head is not None and self.late_bias_norm is not None
                and isinstance(self.late_bias_gate, nn.Parameter)):
            c = ag.layer_norm(ctx, self.late_bias_norm.weight, self.late_bias_norm.bias, self.late_bias_norm.eps)
            late = ag.linear(c, self.late_bias_head.weight, self.late_bias_head.bias)             # [B, N, steps]
            gate = self.late_bias_gate.reshape(-1)
        if isinstance(self.min_sigma_vector, torch.Tensor) and self.min_sigma_vector.numel() > 0:
            floor = self.min_sigma_vector.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        else:
            floor = torch.full((N,), self.min_sigma, dtype=torch.float32, device=dev)
        return ag.nb_head(seq, Wt, bt, self.mu_head.weight, self.mu_head.bias, self.sigma_head.weight, self.sigma_head.bias,
                          hist.contiguous(), late, gate, floor)

    def stack_forward(self, seq: torch.Tensor) -> torch.Tensor:
        """TimesBlock stack on pre-embedded features ``[B, L, d_model]``: ``n_layers x (TimesBlock + residual +
        shared LayerNorm)`` (timesnet.py:2050-2061).  Sync-free, so it can be captured in a CUDA graph."""
        for block in self.blocks:
            object.__setattr__(block, "period_selector", self.period_selector)
            seq = block.forward_norm(seq, self.layer_norm)
        return seq

    def graphed(self, x: torch.Tensor, x_mark: Optional[torch.Tensor] = None,
                series_static: Optional[torch.Tensor] = None, series_ids: Optional[torch.Tensor] = None):
        """CUDA-graph replay of ``forward`` for inputs shaped like the given examples.  Returns a callable
        taking the same tensors (positionally, ``None`` entries dropped).  Turns ``check_finite`` off:
        the two host-syncing sanity checks of timesnet.py:2094-2097 cannot live inside a graph."""
        from ..cuda_graphs import GraphedCallable
        names = ["x", "x_mark", "series_static", "series_ids"]
        given = [x, x_mark, series_static, series_ids]
        idx = [i for i, t in enumerate(given) if t is not None]

        def run(*tensors):
            kw = {names[i]: t for i, t in zip(idx, tensors)}
            keep = self.check_finite
            self.check_finite = False                       # scoped to the graphed callable (capture and warm-up)
            try:
                return self.forward(**kw)
            finally:
                self.check_finite = keep

        run(*[given[i] for i in idx])                       # lazy build outside of any capture
        return GraphedCallable(run, [given[i] for i in idx], params_of=self)

    def forward(self, x: torch.Tensor, x_mark: Optional[torch.Tensor] = None,
                series_static: Optional[torch.Tensor] = None,
                series_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if x.ndim != 3:
            raise ValueError("TimesNet expects input shaped [B, T, N]")
        B, T, N = x.shape
        if T < self.input_len:
            raise ValueError(f"Input sequence length {T} is shorter than required input_len {self.input_len}")
        if x_mark is not None and x_mark.shape[:2] != x.shape[:2]:
            raise ValueError("x_mark must share batch/time dimensions with x")
        nv.require_cuda(x, "x")
        if x.dtype != torch.float32:
            raise TypeError("TimesNet.forward takes float32 input (the reference rejects half inputs too); "
                            "use stack_dtype=torch.bfloat16 for the bf16 TimesBlock stack")
        if _wants_grad(self, x):
            return self._forward_differentiable(x, x_mark, series_static, series_ids)
        _forward_only_guard(self, x, self.dropout)
        L = self.input_len
        dev = x.device
        with torch.no_grad():
            xv = x[:, -L:, :].contiguous()                                   # LAST input_len steps (timesnet.py:1877)
            mark = x_mark[:, -L:, :] if x_mark is not None else None
            self._ensure_embedding(xv, mark, series_static, series_ids)
            steps = self.pred_len if self.mode == "direct" else self._out_steps
            ctx = self._context(B, N, dev, series_static, series_ids)
            feat_in = xv
            if (ctx is not None and self.use_zero_mean_context and self.context_coeff is not None
                    and self.temporal_context is not None):
                coeff = nv.linear(ctx, self._f32(self.context_coeff.weight), self._f32(self.context_coeff.bias))
                feat_in = self.temporal_context.add_to(xv, coeff)            # timesnet.py:1966-1983
            if ctx is not None and self.use_constant_context_bias and self.context_proj is not None:
                cb = nv.linear(ctx, self._f32(self.context_proj.weight), self._f32(self.context_proj.bias))
                feat_in = feat_in + cb.squeeze(-1).unsqueeze(1)              # timesnet.py:1984-1991
            sdt = self.stack_dtype or torch.float32
            seq = self.embedding(feat_in, mark, out_dtype=sdt)               # timesnet.py:1996
            if seq.size(1) != L or seq.size(-1) != self.d_model:
                raise RuntimeError("Embedding output must have shape [B, input_len, d_model]")
            seq = self.stack_forward(seq)                                    # timesnet.py:2050-2061
            # ---- head (timesnet.py:2008-2014, 2063-2093) ----
            hist_steps = min(steps, L)
            hist = xv[:, -hist_steps:, :]                                    # a VIEW of x: the head reads it in place
            if hist_steps < steps:                                           # pred_len > input_len: repeat the last step
                hist = torch.cat([hist, hist[:, -1:, :].expand(-1, steps - hist_steps, -1)], dim=1).contiguous()
            Wt = self._f32(self.forecast_time_proj.weight[-steps:, :] if steps != self.pred_len
                           else self.forecast_time_proj.weight)
            bt = self._f32(self.forecast_time_proj.bias[-steps:] if steps != self.pred_len
                           else self.forecast_time_proj.bias)
            late = gate = None
            if (ctx is not None and self.late_bias_head is not None and self.late_bias_norm is not None
                    and isinstance(self.late_bias_gate, nn.Parameter)):
                late = self._late_bias(ctx)
                gate = self._f32(self.late_bias_gate).reshape(-1)
            if isinstance(self.min_sigma_vector, torch.Tensor) and self.min_sigma_vector.numel() > 0:
                floor = self.min_sigma_vector.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            else:
                fkey = (N, self.min_sigma, str(dev))
                if getattr(self, "_floor_cache", (None,))[0] != fkey:
                    self._floor_cache = (fkey, torch.full((N,), self.min_sigma, dtype=torch.float32, device=dev))
                floor = self._floor_cache[1]
            flags = torch.zeros(1, dtype=torch.int32, device=dev)
            res = None
            if self.d_model % 16 == 0 and N >= 16:
                w_heads, b_heads, n_pad = self._heads_split(dev)
                late_t = self._late_bias_gated_t(late, gate) if late is not None else None
                res = nv.nb_head_tc(seq, steps, N, Wt, bt, w_heads, b_heads, n_pad, hist,
                                    late if late_t is None else late_t, gate if late_t is None else None, floor, flags,
                                    wt_s3=self._time_proj_split(Wt, steps, dev))
            if res is None:
                res = nv.nb_head(seq, steps, N, Wt, bt, self._f32(self.mu_head.weight),
                                 self._f32(self.mu_head.bias), self._f32(self.sigma_head.weight),
                                 self._f32(self.sigma_head.bias), hist, late, gate, floor, flags)
            rate, disp = res
            if self.check_finite:
                bad = int(flags.item())                                      # the reference syncs here too
                if bad & 1:
                    raise RuntimeError("Predicted rate must be finite and strictly positive")
                if bad & 2:
                    raise RuntimeError("Predicted dispersion must be finite and strictly positive")
        return rate, disp
