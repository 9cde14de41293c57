"""CPU oracle for the TimesBlock forward path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional (state-dict in, tensors out) restatement of the
reference's PyTorch algorithm for the hot path named in BASELINE.json.  It is
the checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package never does.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference
from ``/root/reference/src`` (only possible in the build container), runs it
and this file on the same seeded inputs/weights, asserts agreement and writes
the fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks
this file against those fixtures everywhere (including the GPU box, where
``/root/reference`` does not exist).

All reference citations are ``file:line`` under ``/root/reference/src/
timesnet_forecast/`` (``timesnet.py`` = ``models/timesnet.py``).

Everything here runs in torch fp32 on the CPU (the reference's own ground
truth); the float64 numpy DFT variant at the bottom exists to
measure spectral gaps when judging tie-sensitivity of the integer period
selection.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]


# --------------------------------------------------------------------------- #
# a1: shared FFT period selector                     timesnet.py:64-159
# --------------------------------------------------------------------------- #
@dataclass
class Selection:
    freq_indices: Tensor          # [K'] int64, bins that survived the cycle filter
    periods: Tensor               # [K'] int64
    amplitudes: Tensor            # [B, K'] x.dtype, channel-median amplitude at the bins
    raw_indices: Tensor           # [k] int64, top-k order before filtering
    amp_median: Optional[Tensor]  # [B, F] fp32 per-sample channel-median spectrum
    amp_mean: Optional[Tensor]    # [F] fp32 batch mean (before the dtype cast)
    scores: Optional[Tensor]      # [F] x.dtype penalised score that was ranked


def channel_median_spectrum(x: Tensor) -> Tensor:
    """|rfft| over time followed by the lower median over channels.

    timesnet.py:92-111.  Half inputs are upcast to fp32 first (:93-94).
    Returns ``[B, L//2+1]`` fp32.
    """
    xf = x.to(torch.float32) if x.dtype in (torch.float16, torch.bfloat16) else x
    amp = torch.abs(torch.fft.rfft(xf, dim=1))          # [B, F, C]
    return amp.median(dim=2).values                      # lower median (torch semantics)


def select_from_spectrum(
    amp_median: Tensor,
    amp_sum: Tensor,
    global_batch: int,
    seq_len: int,
    k: int,
    pmax: int,
    min_period_threshold: int,
    dtype: torch.dtype,
) -> Selection:
    """Tail of the selector once the batch-summed spectrum is known.

    ``amp_sum`` is the sum over the *global* batch of ``amp_median`` rows (so
    the multi-GPU path can all-reduce it); ``amp_median`` holds the local rows
    used for the per-sample amplitudes.  timesnet.py:112-159.
    """
    L = int(seq_len)
    B = amp_median.shape[0]
    empty_idx = torch.zeros(0, dtype=torch.long)
    empty = Selection(empty_idx, empty_idx, torch.zeros(B, 0, dtype=dtype), empty_idx,
                      amp_median, None, None)
    k_cfg = int(max(0, k))
    pmax = int(max(1, pmax))
    mpt = int(min(pmax, max(1, min_period_threshold)))
    if k_cfg <= 0 or L <= 1 or B <= 0:
        return empty
    amp_mean32 = amp_sum / float(global_batch)
    nbins = amp_mean32.numel()
    if nbins <= 1:
        return empty
    amp_mean = amp_mean32.to(dtype).clone()              # :119 (bf16 rounding happens here)
    amp_mean[0] = float("-inf")                          # :120
    kk = min(k_cfg, nbins - 1)                           # :122-123
    if kk <= 0:
        return empty
    log_idx = torch.log1p(torch.arange(nbins, dtype=torch.long).to(torch.float32))  # :128-129
    scores = amp_mean - 1e-8 * log_idx.to(dtype)          # :130
    _, idx = torch.topk(scores, k=kk, largest=True)      # :131
    safe = idx.to(torch.long).clamp_min(1)               # :132
    sample = amp_median.gather(1, safe.view(1, -1).expand(B, -1))   # :133-135
    upper = min(pmax, max(1, L - 1))                     # :138
    lower = mpt
    if upper < lower:
        return empty
    periods = (L + safe - 1) // safe                     # :144
    periods = torch.clamp(periods, min=lower, max=upper)  # :145
    cycles = (L + periods - 1) // periods                # :147
    valid = cycles >= 2                                  # :148
    if not bool(valid.any()):
        return empty
    return Selection(
        freq_indices=safe[valid],
        periods=periods[valid],
        amplitudes=sample[:, valid].to(dtype),
        raw_indices=safe,
        amp_median=amp_median,
        amp_mean=amp_mean32,
        scores=scores,
    )


def select_periods(x: Tensor, k: int, pmax: int, min_period_threshold: int = 1) -> Selection:
    """Full single-process selector, timesnet.py:64-159."""
    if x.ndim != 3:
        raise ValueError("FFTPeriodSelector expects input shaped [B, L, C]")
    B, L, C = x.shape
    empty_idx = torch.zeros(0, dtype=torch.long)
    if int(max(0, k)) <= 0 or L <= 1 or C <= 0 or B <= 0:
        return Selection(empty_idx, empty_idx, torch.zeros(B, 0, dtype=x.dtype), empty_idx,
                         None, None, None)
    med = channel_median_spectrum(x)
    # ``mean(dim=0)`` in the reference; written as sum / B so the sharded path
    # (sum of per-rank sums) is the same expression.  Verified bit-identical to
    # torch.mean on the golden inputs by make_golden.py.
    return select_from_spectrum(med, med.sum(dim=0), B, L, k, pmax, min_period_threshold, x.dtype)


# --------------------------------------------------------------------------- #
# a2: PeriodGrouper                                  timesnet.py:286-557
# --------------------------------------------------------------------------- #
@dataclass
class Groups:
    periods: List[int] = field(default_factory=list)      # ascending (period, canonical index)
    pads: List[int] = field(default_factory=list)
    cycles: List[int] = field(default_factory=list)
    mapping: List[int] = field(default_factory=list)      # candidate -> group or -1
    canonical: List[int] = field(default_factory=list)
    logits: Optional[Tensor] = None                       # [B, G] logsumexp of member amplitudes


def _log_bucket(p: int, base: float) -> int:
    # timesnet.py:350-354 (fp32 log, +1e-6, floor)
    v = torch.log(torch.tensor(float(p), dtype=torch.float32)) / math.log(base)
    return int(torch.floor(v + 1e-6).item())


def group_periods(
    periods: Sequence[int],
    amplitudes: Tensor,
    seq_len: int,
    min_period: Optional[int] = None,
    max_period: Optional[int] = None,
    log_base: Optional[float] = None,
    max_unique: Optional[int] = None,
) -> Groups:
    """Plain-Python restatement of ``PeriodGrouper.group`` (timesnet.py:513-557).

    ``log_base`` / ``max_unique`` are the resolved values of the
    ``TIMES_PERIOD_BINNING`` / ``TIMES_PERIOD_MAX_UNIQ`` env opt-ins (:320-325);
    ``None`` (the default) merges exact duplicate periods only.
    """
    L = int(seq_len)
    per = [int(p) for p in periods]
    amp = amplitudes if amplitudes.dim() == 2 else amplitudes.view(1, -1)
    K = len(per)
    out = Groups(mapping=[-1] * K, logits=torch.zeros(amp.shape[0], 0, dtype=amp.dtype))
    cand = []
    for i, p in enumerate(per):
        if p <= 0:
            continue                                      # :517
        if min_period is not None and p < min_period:
            continue                                      # :521-522
        if max_period is not None and p > max_period:
            continue                                      # :523-524
        pad = (-L) % p                                    # :531
        cyc = (L + pad) // p                              # :532-533
        if cyc < 2:
            continue                                      # :534
        cand.append((i, p, pad, cyc))
    if not cand:
        return out
    keys = [(_log_bucket(p, log_base) if log_base is not None else p) for (_, p, _, _) in cand]  # :547-550
    uniq = sorted(set(keys))
    assign = [uniq.index(kv) for kv in keys]              # :551 unique(sorted, return_inverse)
    amp_sel = amp[:, [c[0] for c in cand]]                # :545

    def metadata(assign_now):
        info = []
        for gid in sorted(set(assign_now)):               # :365
            members = [j for j, a in enumerate(assign_now) if a == gid]
            logits = torch.logsumexp(amp_sel[:, members], dim=1)          # :373
            if len(members) == 1:
                best = 0
            else:
                best = int(torch.argmax(amp_sel[:, members].mean(dim=0)).item())   # :377
            canon = members[best]
            info.append(dict(id=gid, members=members, canon=canon, period=cand[canon][1],
                             pad=cand[canon][2], cycles=cand[canon][3], logits=logits,
                             score=float(logits.mean().item()), canon_index=cand[canon][0]))
        return info

    if max_unique is not None and len(set(assign)) > max_unique:          # :394-437
        info = metadata(assign)
        score = torch.tensor([it["score"] for it in info], dtype=torch.float32)
        keep = torch.topk(score, k=max_unique, largest=True).indices.tolist()
        keep_periods = torch.tensor([float(info[j]["period"]) for j in keep], dtype=torch.float32)
        new_assign = list(assign)
        for j, it in enumerate(info):
            if j in keep:
                continue
            d = torch.abs(keep_periods - float(it["period"]))
            target = info[keep[int(torch.argmin(d).item())]]["id"]
            for m in it["members"]:
                new_assign[m] = target
        assign = new_assign

    info = metadata(assign)
    info.sort(key=lambda it: (it["period"], it["canon_index"]))           # :453-458
    logits = []
    for g, it in enumerate(info):
        for m in it["members"]:
            out.mapping[cand[m][0]] = g                                   # :476-477
        out.periods.append(it["period"])
        out.pads.append(it["pad"])
        out.cycles.append(it["cycles"])
        out.canonical.append(it["canon_index"])
        logits.append(it["logits"])
    out.logits = torch.stack(logits, dim=1)
    return out


def group_weights(amplitudes: Tensor, mapping: Sequence[int], n_groups: int) -> Tensor:
    """softmax over the valid raw candidates (fp32) scattered into groups.

    timesnet.py:992-1009.  Result dtype = amplitude dtype (bf16 weights are
    rounded *before* the scatter-add, exactly like the reference).
    """
    amp = amplitudes if amplitudes.dim() == 2 else amplitudes.view(1, -1)
    valid = [i for i, g in enumerate(mapping) if g >= 0]
    sm = F.softmax(amp[:, valid].to(torch.float32), dim=1).to(amp.dtype)
    w = torch.zeros(amp.shape[0], n_groups, dtype=amp.dtype)
    idx = torch.tensor([mapping[i] for i in valid], dtype=torch.long).view(1, -1).expand(amp.shape[0], -1)
    w.scatter_add_(1, idx, sm)
    return w


# --------------------------------------------------------------------------- #
# a4: Inception bank                                  timesnet.py:560-654, 731-765
# --------------------------------------------------------------------------- #
def _act(z: Tensor, act: str) -> Tensor:
    return F.relu(z) if act == "relu" else F.gelu(z)      # nn.GELU() default = exact erf (:643)


def count_paths(w: Weights, prefix: str) -> int:
    n = 0
    while f"{prefix}paths.{n}.branch.0.weight" in w:
        n += 1
    return n


def inception_block(x: Tensor, w: Weights, prefix: str, act: str) -> Tensor:
    """``InceptionBlock.forward`` on an NCHW grid (timesnet.py:645-654).

    Branch structure is read off the state dict: ``branch.{0}`` only = single
    k x k conv (ratio 1, :575-580), ``branch.{0,1,2}`` = 1x1 -> k x k -> 1x1
    bottleneck with no activation in between (:586-590).  Every conv zero-pads
    its own input by k//2 (:574).
    """
    feats = []
    for j in range(count_paths(w, prefix)):
        h = x
        i = 0
        while f"{prefix}paths.{j}.branch.{i}.weight" in w:
            wt = w[f"{prefix}paths.{j}.branch.{i}.weight"]
            bs = w[f"{prefix}paths.{j}.branch.{i}.bias"]
            h = F.conv2d(h, wt, bs, padding=(wt.shape[2] // 2, wt.shape[3] // 2))
            i += 1
        feats.append(h)
    z = torch.cat(feats, dim=1)                                           # :650
    z = F.conv2d(z, w[f"{prefix}proj.weight"], w[f"{prefix}proj.bias"])   # :651
    z = _act(z, act)                                                      # :652 (dropout = eval identity)
    if f"{prefix}res_proj.weight" in w:
        res = F.conv2d(x, w[f"{prefix}res_proj.weight"], w[f"{prefix}res_proj.bias"])   # :648
    else:
        res = x                                                           # Identity when in == out (:637)
    return z + res                                                        # :654


def inception_stack(x: Tensor, w: Weights, prefix: str, act: str) -> Tensor:
    """``Sequential(InceptionBlock, act, InceptionBlock)`` (timesnet.py:744-762)."""
    h = inception_block(x, w, f"{prefix}0.", act)
    h = _act(h, act)
    return inception_block(h, w, f"{prefix}2.", act)


# --------------------------------------------------------------------------- #
# a3 + a5: TimesBlock                                timesnet.py:767-818, 955-1101
# --------------------------------------------------------------------------- #
@dataclass
class BlockTrace:
    out: Tensor
    groups: Optional[Groups] = None
    weights: Optional[Tensor] = None          # [B, G]
    deltas: Optional[List[Tensor]] = None     # G x [B, L, C] in x.dtype
    selection: Optional[Selection] = None


def period_delta(x: Tensor, period: int, pad: int, cycles: int, w: Weights, prefix: str, act: str) -> Tensor:
    """Fold -> inception -> minus grid -> unfold for one period (timesnet.py:1037-1070).

    Convs run in fp32 with fp32 weights even for half activations (:1047-1052,
    ``TIMES_MP_CONV`` unset); the delta is cast back to x.dtype (:1068-1069).
    """
    B, L, C = x.shape
    xp = x.permute(0, 2, 1)                                   # [B, C, L]  (:966)
    xp = F.pad(xp, (0, pad)) if pad > 0 else xp               # zero tail (:1017)
    grid = xp.reshape(B, C, cycles, period).to(torch.float32)  # :1042, :1050-1052
    conv = inception_stack(grid, w, prefix, act)              # :1056
    delta = (conv - grid).reshape(B, C, cycles * period)[..., :L]   # :1063-1066
    return delta.permute(0, 2, 1).contiguous().to(x.dtype)    # :1067-1069


def timesblock_from_periods(
    x: Tensor,
    periods: Sequence[int],
    amplitudes: Tensor,
    w: Weights,
    prefix: str,
    act: str = "gelu",
    min_period: Optional[int] = None,
    max_period: Optional[int] = None,
    log_base: Optional[float] = None,
    max_unique: Optional[int] = None,
) -> BlockTrace:
    """TimesBlock body given selector output (timesnet.py:955-1101, :818)."""
    B, L, C = x.shape
    if len(periods) == 0:
        return BlockTrace(out=x)                              # :796-797
    amp = amplitudes.to(x.dtype)                              # :803
    g = group_periods(periods, amp, L, min_period, max_period, log_base, max_unique)
    if not g.periods:
        return BlockTrace(out=x, groups=g)                    # :989-990
    wts = group_weights(amp, g.mapping, len(g.periods))       # :999-1009
    deltas = [period_delta(x, p, pd, cy, w, prefix, act) for p, pd, cy in zip(g.periods, g.pads, g.cycles)]
    stacked = torch.stack(deltas, dim=-1)                     # :1075
    combined = (stacked * wts.to(stacked.dtype).view(B, 1, 1, -1)).sum(dim=-1)   # :1076-1092
    return BlockTrace(out=x + combined, groups=g, weights=wts, deltas=deltas)    # :818


def timesblock_forward(
    x: Tensor, w: Weights, prefix: str, k: int, pmax: int, min_period_threshold: int = 1, act: str = "gelu"
) -> BlockTrace:
    """``TimesBlock.forward`` with the shared FFT selector (timesnet.py:767-818)."""
    sel = select_periods(x, k, pmax, min_period_threshold)
    mpt = int(min(max(1, pmax), max(1, min_period_threshold)))
    tr = timesblock_from_periods(x, sel.periods.tolist(), sel.amplitudes, w, prefix, act,
                                 min_period=mpt, max_period=int(max(1, pmax)))
    tr.selection = sel
    return tr


# --------------------------------------------------------------------------- #
# a6: block loop + shared LayerNorm                  timesnet.py:2050-2061, 1162-1181
# --------------------------------------------------------------------------- #
def layer_norm_fp32(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    xc = x.to(torch.float32) if x.dtype in (torch.float16, torch.bfloat16) else x
    y = F.layer_norm(xc, (x.shape[-1],), weight.to(xc.dtype), bias.to(xc.dtype), eps)
    return y.to(x.dtype)


def stack_forward(
    features: Tensor, w: Weights, n_layers: int, k: int, pmax: int, min_period_threshold: int = 1,
    act: str = "gelu", trace: Optional[List[BlockTrace]] = None,
) -> Tensor:
    seq = features
    for i in range(n_layers):
        tr = timesblock_forward(seq, w, f"blocks.{i}.inception.", k, pmax, min_period_threshold, act)
        if trace is not None:
            trace.append(tr)
        delta = tr.out - seq                                  # :2059
        seq = seq + delta                                     # :2060 (eval: dropout = identity)
        seq = layer_norm_fp32(seq, w["layer_norm.weight"], w["layer_norm.bias"])   # :2061
    return seq


# --------------------------------------------------------------------------- #
# a7: LowRankTemporalContext                         timesnet.py:1340-1371
# --------------------------------------------------------------------------- #
def lowrank_basis(length: int, rank: int, dtype: torch.dtype = torch.float32) -> Tensor:
    calc = torch.float32 if dtype in (torch.float16, torch.bfloat16) else dtype
    steps = torch.arange(length, dtype=calc).unsqueeze(1)
    freqs = torch.arange(1, rank + 1, dtype=calc).unsqueeze(0)
    basis = torch.cos(math.pi / float(length) * (steps + 0.5) * freqs)   # :1346
    basis = basis - basis.mean(dim=0, keepdim=True)                     # :1347
    norm = torch.linalg.norm(basis, dim=0, keepdim=True)                # :1348
    basis = basis / norm.clamp_min(torch.finfo(basis.dtype).eps)        # :1349-1350
    return basis.to(dtype)


def lowrank_context(coeff: Tensor, length: int, scale: Tensor) -> Tensor:
    basis = lowrank_basis(length, coeff.shape[-1], coeff.dtype)
    ctx = torch.einsum("lr,bnr->bln", basis, coeff)          # :1368
    ctx = ctx - ctx.mean(dim=1, keepdim=True)                # :1369
    return ctx * scale.to(coeff.dtype)                       # :1370-1371


# --------------------------------------------------------------------------- #
# a10: DataEmbedding                                  timesnet.py:1104-1129, 1257-1325
# --------------------------------------------------------------------------- #
def positional_table(L: int, d_model: int) -> Tensor:
    pos = torch.arange(L, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(L, d_model, dtype=torch.float32)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div[: pe[:, 1::2].shape[1]])
    return pe


def data_embedding(x: Tensor, w: Weights, x_mark: Optional[Tensor] = None, mode: str = "decoupled") -> Tensor:
    value = F.linear(x, w["embedding.value_embedding.weight"], w["embedding.value_embedding.bias"])  # :1295
    d_model = value.shape[-1]
    aux = positional_table(x.shape[1], d_model).to(x.dtype).unsqueeze(0).expand(x.shape[0], -1, -1)  # :1296
    if x_mark is not None and "embedding.temporal_embedding.weight" in w:
        aux = aux + F.linear(x_mark, w["embedding.temporal_embedding.weight"],
                             w["embedding.temporal_embedding.bias"])       # :1297-1302
    if mode == "decoupled":
        auxn = layer_norm_fp32(aux, w["embedding.aux_norm.weight"], w["embedding.aux_norm.bias"])   # :1308
        return value + w["embedding.gate"].to(value.dtype) * auxn          # :1312
    out = value + aux                                                      # :1314
    if mode == "layer":
        return layer_norm_fp32(out, w["embedding.norm.weight"], w["embedding.norm.bias"])   # :1315-1317
    if mode == "rms":
        return rms_norm(out, w["embedding.norm.weight"], w["embedding.norm.bias"])          # :1318-1320
    if mode == "none":
        return out
    raise NotImplementedError(mode)


def rms_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """RMSNorm.forward (timesnet.py:1132-1159)."""
    xc = x.to(torch.float32) if x.dtype in (torch.float16, torch.bfloat16) else x
    scale = torch.rsqrt(xc.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (xc * scale * weight.to(xc.dtype) + bias.to(xc.dtype)).to(x.dtype)


# --------------------------------------------------------------------------- #
# a8 + whole model                                    timesnet.py:1857-2102
# --------------------------------------------------------------------------- #
@dataclass
class ModelCfg:
    input_len: int
    pred_len: int
    d_model: int
    n_layers: int
    k_periods: int
    mode: str = "direct"
    activation: str = "gelu"
    min_period_threshold: int = 1
    min_sigma: float = 1e-3
    use_zero_mean_context: bool = False
    context_rank: int = 0
    static_layernorm: bool = True
    embed_norm_mode: str = "decoupled"


def softplus32(z: Tensor) -> Tensor:
    return F.softplus(z.to(torch.float32), beta=1.0, threshold=20).to(z.dtype)


def context_vector(w: Weights, cfg: ModelCfg, B: int, N: int,
                   series_static: Optional[Tensor], series_ids: Optional[Tensor]) -> Optional[Tensor]:
    """Static projection + id embedding + context LayerNorm (timesnet.py:1886-1957)."""
    comps = []
    if series_static is not None and "static_proj.weight" in w:
        st = series_static.unsqueeze(0).expand(B, -1, -1) if series_static.ndim == 2 else series_static
        sp = F.linear(st, w["static_proj.weight"], w["static_proj.bias"])
        if cfg.static_layernorm and "static_norm.weight" in w:
            sp = layer_norm_fp32(sp, w["static_norm.weight"], w["static_norm.bias"])
        comps.append(sp)
    if "series_embedding.weight" in w:
        ids = torch.arange(N, dtype=torch.long) if series_ids is None else series_ids.to(torch.long)
        ids = ids.unsqueeze(0) if ids.ndim == 1 else ids
        ids = ids.expand(B, -1) if ids.shape[0] == 1 and B > 1 else ids
        comps.append(F.embedding(ids, w["series_embedding.weight"]))
    if not comps:
        return None
    ctx = torch.cat(comps, dim=-1)
    if "context_norm.weight" in w:
        ctx = layer_norm_fp32(ctx, w["context_norm.weight"], w["context_norm.bias"])
    return ctx


def nb_head(seq: Tensor, x_value: Tensor, ctx: Optional[Tensor], w: Weights, cfg: ModelCfg,
            min_sigma_vector: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Time projection + mu/sigma heads + softplus/floor (timesnet.py:2008-2014, 2063-2093)."""
    target = cfg.pred_len if cfg.mode == "direct" else 1
    L = x_value.shape[1]
    hist = min(target, L)
    tail = x_value[:, -hist:, :]
    if hist < target:
        tail = torch.cat([tail, tail[:, -1:, :].expand(-1, target - hist, -1)], dim=1)
    fb = seq.permute(0, 2, 1)
    full = F.linear(fb, w["forecast_time_proj.weight"], w["forecast_time_proj.bias"])    # :2071
    base = full[:, :, -target:] if target != cfg.pred_len else full                       # :2072-2075
    hidden = base.permute(0, 2, 1)
    pre = F.linear(hidden, w["mu_head.weight"], w["mu_head.bias"]) + tail                 # :2079
    if ctx is not None and "late_bias_head.weight" in w:                                  # :2028-2048
        c = layer_norm_fp32(ctx, w["late_bias_norm.weight"], w["late_bias_norm.bias"])
        bias = F.linear(c, w["late_bias_head.weight"], w["late_bias_head.bias"]).permute(0, 2, 1)
        pre = pre + w["late_bias_gate"].to(pre.dtype) * bias
    rate = softplus32(pre) + 1e-6                                                         # :2081-2085
    sig = softplus32(F.linear(hidden, w["sigma_head.weight"], w["sigma_head.bias"]))      # :2087-2091
    if min_sigma_vector is not None and min_sigma_vector.numel() > 0:
        floor = min_sigma_vector.reshape(1, 1, -1).to(sig.dtype).expand_as(sig)           # :1851-1854
    else:
        floor = torch.full_like(sig, cfg.min_sigma)
    return rate, sig + floor + 1e-6                                                       # :2092-2093


def timesnet_forward(
    x: Tensor, w: Weights, cfg: ModelCfg, x_mark: Optional[Tensor] = None,
    series_static: Optional[Tensor] = None, series_ids: Optional[Tensor] = None,
    min_sigma_vector: Optional[Tensor] = None, trace: Optional[dict] = None,
) -> Tuple[Tensor, Tensor]:
    """``TimesNet.forward`` in eval mode (timesnet.py:1857-2102)."""
    B, T, N = x.shape
    L = cfg.input_len
    xv = x[:, -L:, :]                                                    # :1877 (last input_len steps)
    mark = x_mark[:, -L:, :] if x_mark is not None else None
    feat_in = xv.clone()
    ctx = context_vector(w, cfg, B, N, series_static, series_ids)
    if ctx is not None and cfg.use_zero_mean_context and "context_coeff.weight" in w:
        coeff = F.linear(ctx, w["context_coeff.weight"], w["context_coeff.bias"])        # :1966
        feat_in = feat_in + lowrank_context(coeff, L, w["temporal_context.scale"])       # :1967-1983
    if ctx is not None and "context_proj.weight" in w:
        feat_in = feat_in + F.linear(ctx, w["context_proj.weight"], w["context_proj.bias"]).squeeze(-1).unsqueeze(1)
    features = data_embedding(feat_in, w, mark, cfg.embed_norm_mode)                     # :1996
    blocks: List[BlockTrace] = []
    seq = stack_forward(features, w, cfg.n_layers, cfg.k_periods, L, cfg.min_period_threshold,
                        cfg.activation, trace=blocks)
    rate, disp = nb_head(seq, xv, ctx, w, cfg, min_sigma_vector)
    if trace is not None:
        trace.update(features=features, blocks=blocks, seq=seq)
    return rate, disp


# --------------------------------------------------------------------------- #
# a9: NB negative log-likelihood                      losses.py:6-58
# --------------------------------------------------------------------------- #
def nb_mask(y: Tensor, rate: Tensor, disp: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    m = torch.isfinite(y) & torch.isfinite(rate) & torch.isfinite(disp)
    if mask is not None:
        mb = mask.to(torch.bool)
        if mb.ndim < m.ndim:
            mb = mb.reshape(*mb.shape, *([1] * (m.ndim - mb.ndim)))
        m = m & mb.expand_as(m)
    return m


def nb_nll(y: Tensor, rate: Tensor, disp: Tensor, mask: Optional[Tensor] = None, eps: float = 1e-8) -> Tensor:
    y = torch.clamp(y.to(torch.float32), min=0.0)
    alpha = torch.clamp(disp.to(torch.float32), min=eps)
    mu = torch.clamp(rate.to(torch.float32), min=eps)
    l1p = torch.log1p(alpha * mu)
    inv = torch.reciprocal(alpha)
    ll = (torch.lgamma(y + inv) - torch.lgamma(inv) - torch.lgamma(y + 1.0)
          + inv * (-l1p) + y * (torch.log(alpha) + torch.log(mu) - l1p))
    wgt = nb_mask(y, mu, alpha, mask).to(torch.float32)
    return -(ll * wgt).sum() / torch.clamp(wgt.sum(), min=1.0)


# --------------------------------------------------------------------------- #
# a11: recursive rolling forecast                     predict.py:307-342
# --------------------------------------------------------------------------- #
def forecast_recursive(x: Tensor, H: int, w: Weights, cfg: ModelCfg, **kw) -> Tuple[Tensor, Tensor]:
    assert cfg.mode == "recursive"
    seq = x
    rates, disps = [], []
    for _ in range(H):
        r, d = timesnet_forward(seq, w, cfg, **kw)
        rates.append(r)
        disps.append(d)
        seq = torch.cat([seq[:, 1:, :], r], dim=1)
    return torch.cat(rates, dim=1), torch.cat(disps, dim=1)


# --------------------------------------------------------------------------- #
# float64 cross-check of the integer selection (tie / gap diagnostics)
# --------------------------------------------------------------------------- #
def spectrum_float64(x: Tensor) -> np.ndarray:
    """Batch-mean channel-median spectrum in float64 numpy (direct rfft)."""
    a = np.abs(np.fft.rfft(x.to(torch.float64).numpy(), axis=1))        # [B, F, C]
    C = a.shape[2]
    med = np.sort(a, axis=2)[:, :, (C - 1) // 2]                          # lower median
    return med.mean(axis=0)


def score_gap(amp_mean: np.ndarray, k: int) -> float:
    """Relative gap between the k-th and (k+1)-th ranked bins (DC excluded)."""
    s = np.sort(amp_mean[1:])[::-1]
    if s.size <= k:
        return float("inf")
    return float((s[k - 1] - s[k]) / max(abs(s[k - 1]), 1e-30))
