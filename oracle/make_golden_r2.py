"""Round-2 golden fixtures (same protocol as make_golden.py: run the REAL reference from
/root/reference/src, check the oracle against it, store the reference's outputs).

    python oracle/make_golden_r2.py            # writes tests/golden/r2_*.pt

Cases ("next" item 1):
  r2_stack_traffic.pt   traffic-shaped TimesBlock stack (L=720, C=256, F=1024, mid=64, 3 layers), B=2, f32 + bf16
  r2_recursive5.pt      BASELINE config 5 shape (N=1, L=28, k=2, mpt=7, R=16, statics) at B=512: one forward + the
                        28-step rolling forecast
  r2_embed_mark.pt      DataEmbedding with time marks in every norm mode (+ 4-D input), TimesNet.forward with x_mark,
                        recursive forecast with y_mark
  r2_block_env.pt       TimesBlock under TIMES_PERIOD_MAX_UNIQ / TIMES_PERIOD_BINNING (tests/test_times_block.py:183-211)
  r2_inception_nchw.pt  InceptionBlock / InceptionBranch forward on NCHW grids (tests/test_inception_block.py)
Only outputs (sub-sampled where large) are stored; inputs and weights regenerate from seeds.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
REF_SRC = Path("/root/reference/src")
if not REF_SRC.exists():
    raise SystemExit("make_golden_r2.py needs /root/reference/src (build container only)")
sys.path.insert(0, str(REF_SRC))

import flowtimes_oracle as orc                      # noqa: E402
import flowtimes_synth as syn                       # noqa: E402
from timesnet_forecast.models import timesnet as ref  # noqa: E402  (the real reference)
from timesnet_forecast.predict import forecast_recursive_batch as ref_recursive  # noqa: E402
from make_golden import FixedSelector, check, subsample, _ref_block, _ref_stack  # noqa: E402

OUT = ROOT / "tests" / "golden"
torch.set_num_threads(max(1, os.cpu_count() or 1))


def stack_traffic():
    out = {}
    wl = syn.Workload(**{**syn.WORKLOADS["traffic"].__dict__, "B": 2})
    w = syn.stack_weights(wl, seed=0)
    for kind in ("planted", "white"):
        x = (syn.planted_features(wl.B, wl.T, wl.d_model, 0) if kind == "planted"
             else syn.white_features(wl.B, wl.T, wl.d_model, 1))
        for dname in ("f32", "bf16"):
            xd = x.to(syn.torch_dtype(dname))
            y_ref, periods, outs = _ref_stack(wl, w, xd)
            trace = []
            y_orc = orc.stack_forward(xd, w, wl.n_layers, wl.k_periods, wl.T, wl.min_period_threshold, trace=trace)
            d = check(f"traffic.{kind}.{dname}", y_orc, y_ref, 1e-5 if dname == "f32" else 1.6e-2)
            out[f"traffic.{kind}.{dname}"] = dict(
                workload=wl.as_dict(), weight_seed=0, input=kind, periods=[p.tolist() for p in periods],
                block0_weights=trace[0].weights.float(), out_full=None, out_sub=subsample(y_ref),
                block0_out_sub=subsample(outs[0]), out_abs_mean=y_ref.float().abs().mean().item(), oracle_vs_ref=d)
            print("stack", kind, dname, [p.tolist() for p in periods], f"oracle-ref {d:.1e}", flush=True)
    torch.save(out, OUT / "r2_stack_traffic.pt")


def _ref_model(wl, x, static, ids, x_mark=None, embed_norm_mode=None, seed=9):
    torch.manual_seed(0)
    m = ref.TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods,
                     kernel_set=[list(k) for k in wl.kernel_set], dropout=0.1, activation="gelu", mode=wl.mode,
                     d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio, min_period_threshold=wl.min_period_threshold,
                     use_checkpoint=False, use_zero_mean_context=wl.context_rank > 0, context_rank=wl.context_rank,
                     context_scale=0.05, embed_norm_mode=embed_norm_mode)
    kw = {}
    if x_mark is not None:
        kw["x_mark"] = x_mark[:1]
    m(x[:1], series_static=None if static is None else (static[:1] if static.ndim == 3 else static),
      series_ids=ids, **kw)
    m.eval()
    sd = syn.reseed_module_state(m, seed=seed)
    m.load_state_dict(sd, strict=True)
    return m, sd


def recursive5():
    base = syn.WORKLOADS["recursive"]
    wl = syn.Workload(**{**base.__dict__, "B": 512})
    g = torch.Generator().manual_seed(0)
    x = torch.poisson(torch.full((wl.B, wl.T, wl.N), 4.0), generator=g)
    static = torch.randn(wl.B, wl.N, wl.static_features, generator=g)
    ids = torch.arange(wl.N)
    m, sd = _ref_model(wl, x, static, ids)
    cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu", wl.min_period_threshold, 1e-3,
                       True, wl.context_rank)
    r_ref, d_ref = m(x, series_static=static, series_ids=ids)
    r_orc, d_orc = orc.timesnet_forward(x, sd, cfg, series_static=static, series_ids=ids)
    check("recursive5.rate", r_orc, r_ref)
    check("recursive5.disp", d_orc, d_ref)
    periods = m.period_selector.last_selected_periods.tolist()
    rr, rd = ref_recursive(m, x, wl.H, series_static=static, series_ids=ids)
    orr, ord_ = orc.forecast_recursive(x, 3, sd, cfg, series_static=static, series_ids=ids)
    check("recursive5.rec_rate[:3]", orr, rr[:, :3], 2e-5)
    check("recursive5.rec_disp[:3]", ord_, rd[:, :3], 2e-5)
    torch.save(dict(workload=wl.as_dict(), x_seed=0, state_seed=9, state_keys={k: tuple(v.shape) for k, v in sd.items()},
                    rate=r_ref, disp=d_ref, last_periods=periods, rec_rate=rr, rec_disp=rd), OUT / "r2_recursive5.pt")
    print("recursive5: last-layer periods", periods, "rate mean", float(r_ref.mean()), flush=True)


def embed_mark():
    out = {}
    g = torch.Generator().manual_seed(31)
    B, L, N, C, Tm = 3, 24, 5, 16, 4
    x = torch.randn(B, L, N, generator=g)
    mark = torch.randn(B, L, Tm, generator=g)
    for mode in ("decoupled", "none", "layer", "rms"):
        torch.manual_seed(1)
        emb = ref.DataEmbedding(N, C, dropout=0.0, time_features=Tm, embed_norm_mode=mode).eval()
        sd = syn.reseed_module_state(emb, seed=5)
        emb.load_state_dict(sd, strict=True)
        y = emb(x, mark)
        y0 = emb(x)                                                    # marks are optional at call time
        w = {"embedding." + k: v for k, v in sd.items()}
        check(f"embed.{mode}", orc.data_embedding(x, w, mark, mode), y)
        check(f"embed.{mode}.nomark", orc.data_embedding(x, w, None, mode), y0)
        out[f"embed.{mode}"] = dict(x=x, mark=mark, state=sd, out=y, out_nomark=y0, mode=mode)
    # 4-D input [B, L, N, C] with 3-D marks (timesnet.py:1266-1288)
    torch.manual_seed(1)
    emb = ref.DataEmbedding(3, C, dropout=0.0, time_features=Tm).eval()
    sd = syn.reseed_module_state(emb, seed=6)
    emb.load_state_dict(sd, strict=True)
    x4 = torch.randn(2, L, 4, 3, generator=g)
    out["embed.4d"] = dict(x=x4, mark=mark[:2], state=sd, out=emb(x4, mark[:2]))
    # whole model with time marks, direct and recursive (y_mark feeds the rolling marks, predict.py:336-341)
    for name, wl, T_in in (
            ("mark_direct", syn.Workload("mark_direct", 3, 48, 5, 12, 16, 2, 3, "f32", d_ff=32), 48),
            ("mark_recursive", syn.Workload("mark_recursive", 4, 28, 2, 5, 16, 2, 2, "f32", d_ff=32, min_period_threshold=7,
                                            mode="recursive", context_rank=4), 28)):
        x = syn.planted_series(wl.B, T_in, wl.N, seed=3)
        xm = torch.randn(wl.B, T_in, Tm, generator=g)
        ym = torch.randn(wl.B, wl.H, Tm, generator=g)
        ids = torch.arange(wl.N)
        m, sd = _ref_model(wl, x, None, ids, x_mark=xm)
        cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu", wl.min_period_threshold,
                           1e-3, wl.context_rank > 0, wl.context_rank)
        r, d = m(x, x_mark=xm, series_ids=ids)
        ro, do = orc.timesnet_forward(x, sd, cfg, x_mark=xm, series_ids=ids)
        check(name + ".rate", ro, r)
        check(name + ".disp", do, d)
        rec = {}
        if wl.mode == "recursive":
            rr, rd = ref_recursive(m, x, wl.H, x_mark=xm, y_mark=ym, series_ids=ids)
            rec = dict(rec_rate=rr, rec_disp=rd)
        out[name] = dict(workload=wl.as_dict(), T_in=T_in, x_seed=3, x_mark=xm, y_mark=ym, state=sd, rate=r, disp=d, **rec)
    torch.save(out, OUT / "r2_embed_mark.pt")
    print("embed_mark:", list(out), flush=True)


def block_env():
    out = {}
    wl = syn.WORKLOADS["toy"]
    w = syn.stack_weights(wl, seed=3)
    x = syn.white_features(wl.B, wl.T, wl.d_model, seed=4)
    g = torch.Generator().manual_seed(8)
    per = [3, 4, 6, 12, 5]
    amp = torch.randn(wl.B, len(per), generator=g)
    for name, env in (("maxuniq2", {"TIMES_PERIOD_MAX_UNIQ": "2"}),
                      ("log2", {"TIMES_PERIOD_BINNING": "log:2"}),
                      ("log2_maxuniq2", {"TIMES_PERIOD_BINNING": "log:2", "TIMES_PERIOD_MAX_UNIQ": "2"}),
                      ("scheduled", {"TIMES_PERIOD_MAX_UNIQ": "0:3,1:1"})):
        for k_, v_ in env.items():
            os.environ[k_] = v_
        try:
            res = {}
            for depth in (0, 1):
                blk = _ref_block(wl, w, 0, "gelu")
                blk.block_index = depth
                object.__setattr__(blk, "period_selector", FixedSelector(per, amp))
                y = blk(x)
                n_fixed = int(blk._last_group_count)
                base = ref._resolve_log_binning_base(env.get("TIMES_PERIOD_BINNING"), depth)
                mu = ref._resolve_scheduled_int(env.get("TIMES_PERIOD_MAX_UNIQ"), depth)
                tr = orc.timesblock_from_periods(x, per, amp, w, "blocks.0.inception.", "gelu", log_base=base, max_unique=mu)
                check(f"block_env.{name}.{depth}", tr.out, y)
                sel = ref.FFTPeriodSelector(5, wl.T, 1)
                object.__setattr__(blk, "period_selector", sel)
                yf = blk(x)
                res[depth] = dict(fixed_out=y, groups=n_fixed, group_periods=tr.groups.periods,
                                  fft_out=yf, fft_groups=int(blk._last_group_count),
                                  fft_periods=sel.last_selected_periods.tolist())
            out[name] = dict(env=env, workload=wl.as_dict(), weight_seed=3, x=x, periods=per, amps=amp, by_depth=res)
        finally:
            for k_ in env:
                os.environ.pop(k_)
    torch.save(out, OUT / "r2_block_env.pt")
    print("block_env:", {k: {d: (v["by_depth"][d]["groups"], v["by_depth"][d]["fft_groups"]) for d in (0, 1)}
                         for k, v in out.items()}, flush=True)


def inception_nchw():
    out = {}
    g = torch.Generator().manual_seed(17)
    for name, (cin, cout, ratio, ks, shape) in {
            "block_8_8_r0.5": (8, 8, 0.5, [(3, 3), (5, 1)], (2, 8, 5, 7)),
            "block_4_6_r2": (4, 6, 2.0, [(3, 3), (5, 1)], (2, 4, 5, 7)),
            "block_16_32_r4": (16, 32, 4.0, [(3, 3), (5, 5), (7, 7)], (3, 16, 6, 11)),
            "block_8_8_r1": (8, 8, 1.0, [(3, 3), (5, 5)], (2, 8, 4, 9)),
            "block_h1": (8, 12, 2.0, [(3, 3), (1, 5)], (2, 8, 1, 13))}.items():
        torch.manual_seed(2)
        blk = ref.InceptionBlock(cin, cout, ks, dropout=0.0, act="gelu" if "r1" not in name else "relu",
                                 bottleneck_ratio=ratio).eval()
        x = torch.randn(*shape, generator=g)
        out[name] = dict(kind="block", cin=cin, cout=cout, ratio=ratio, kernel_set=ks,
                         act="gelu" if "r1" not in name else "relu",
                         state={k: v.clone() for k, v in blk.state_dict().items()}, x=x, out=blk(x))
    for name, (cin, cout, ratio, k, shape) in {
            "branch_r1_3x5": (4, 6, 1.0, (3, 5), (2, 4, 7, 9)),
            "branch_r2.5_3x3": (3, 7, 2.5, (3, 3), (2, 3, 5, 5)),
            "branch_r4_7x7": (16, 16, 4.0, (7, 7), (2, 16, 4, 6))}.items():
        torch.manual_seed(3)
        br = ref.InceptionBranch(cin, cout, k, ratio).eval()
        x = torch.randn(*shape, generator=g)
        out[name] = dict(kind="branch", cin=cin, cout=cout, ratio=ratio, kernel=k,
                         state={k_: v.clone() for k_, v in br.state_dict().items()}, x=x, out=br(x))
    torch.manual_seed(4)
    rn = ref.RMSNorm(12)
    sd = syn.reseed_module_state(rn, seed=2)
    rn.load_state_dict(sd)
    x = torch.randn(5, 7, 12, generator=g)
    out["rmsnorm"] = dict(kind="rms", state=sd, x=x, out=rn(x), out_bf16=rn(x.bfloat16()).float())
    torch.save(out, OUT / "r2_inception_nchw.pt")
    print("inception_nchw:", list(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["inception_nchw", "block_env", "embed_mark", "stack_traffic", "recursive5"]
    with torch.no_grad():
        for name in which:
            globals()[name]()
    print({p.name: p.stat().st_size for p in OUT.glob("r2_*.pt")})
