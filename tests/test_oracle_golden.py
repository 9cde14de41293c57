"""CPU: the oracle restatement must reproduce the REFERENCE's stored outputs.

The fixtures under tests/golden/ were produced by oracle/make_golden.py from the
unmodified reference (imported from /root/reference/src in the build container).
Here the oracle is re-run on the same seeded inputs and compared with them, so
the pin holds wherever the tests run (the GPU box has no /root/reference).
"""
import torch
import pytest

import flowtimes_oracle as orc
import flowtimes_synth as syn


def _rel(got, want):
    scale = max(1.0, want.float().abs().max().item())
    return (got.float() - want.float()).abs().max().item() / scale


def _sub(t, n=4096):
    flat = t.reshape(-1).float()
    return flat[:: max(1, flat.numel() // n)]


def test_selector_known_answers(golden_dir):
    g = torch.load(golden_dir / "selector_small.pt")
    assert g["ref_shared_L256"]["periods"].tolist() == [64, 32]      # reference tests/test_fft_period_selector.py:14-40
    assert g["ref_bounds_L64"]["periods"].tolist() == [16, 5]        # :43-57
    assert g["ref_zero_k"]["periods"].numel() == 0                   # :60-70
    for name, c in g.items():
        o = orc.select_periods(c["x"], c["k"], c["pmax"], c["mpt"])
        assert o.periods.tolist() == c["periods"].tolist(), name
        assert o.freq_indices.tolist() == c["freq"].tolist(), name
        assert torch.equal(o.amplitudes, c["amps"]), name


@pytest.mark.parametrize("wname", ["etth1", "elec", "traffic", "recursive"])
def test_selector_baseline_shapes(golden_dir, wname):
    g = torch.load(golden_dir / "selector_baseline_shapes.pt")
    for key, c in g.items():
        if not key.startswith(wname + "."):
            continue
        _, kind, dname = key.split(".")
        x = (syn.planted_features(c["B"], c["L"], c["C"], 0) if kind == "planted"
             else syn.white_features(c["B"], c["L"], c["C"], 1)).to(syn.torch_dtype(dname))
        o = orc.select_periods(x, c["k"], c["L"], c["mpt"])
        assert o.periods.tolist() == c["periods"].tolist(), key
        assert o.freq_indices.tolist() == c["freq"].tolist(), key
        assert _rel(o.amplitudes, c["amps"]) == 0.0, key


def test_grouper(golden_dir):
    g = torch.load(golden_dir / "grouper.pt")
    for name, c in g.items():
        base = 2.0 if "TIMES_PERIOD_BINNING" in c["env"] else None
        mu = int(c["env"]["TIMES_PERIOD_MAX_UNIQ"]) if "TIMES_PERIOD_MAX_UNIQ" in c["env"] else None
        o = orc.group_periods(c["periods_in"], c["amps"], c["L"], c["lo"], c["hi"], base, mu)
        assert o.periods == c["periods"].tolist(), name
        assert o.pads == c["pads"].tolist(), name
        assert o.cycles == c["cycles"].tolist(), name
        assert o.mapping == c["mapping"].tolist(), name
        assert o.canonical == c["canonical"].tolist(), name
        if c["logits"].numel():
            assert _rel(o.logits, c["logits"]) < 1e-6, name
    # reference tests/test_times_block.py:157-180 and test_timesblock_vectorized.py:111-129
    assert g["dups_4448"]["periods"].tolist() == [4, 8]
    assert g["too_long"]["periods"].tolist() == [4]                  # period 64 > L=16 is dropped (:88-108)


def test_grouper_weight_mass():
    # reference tests/test_timesblock_vectorized.py:132-167: scatter of softmax == softmax of logsumexp logits
    amps = torch.tensor([[0.5, -0.2, 1.0, -1.5], [1.3, 0.1, -0.4, -2.0]])
    o = orc.group_periods([3, 4, 6, 12], amps, 48, 1, 48, log_base=2.0, max_unique=2)
    w = orc.group_weights(amps, o.mapping, len(o.periods))
    assert len(o.periods) <= 2
    assert torch.allclose(w, torch.softmax(o.logits, dim=1), atol=1e-6)
    assert torch.allclose(w.sum(1), torch.ones(2))


def test_block_toy(golden_dir):
    g = torch.load(golden_dir / "block_toy.pt")
    for name, c in g.items():
        act = name.split(".")[1]
        wd = dict(c["workload"])
        wd["kernel_set"] = tuple(tuple(k) for k in wd["kernel_set"])
        wl = syn.Workload(**wd)
        w = syn.stack_weights(wl, seed=c["weight_seed"])
        y = orc.inception_stack(c["grid"], w, "blocks.0.inception.", act)
        assert _rel(y, c["inception_out"]) < 1e-5, name
        tr = orc.timesblock_from_periods(c["x"], c["fixed_periods"], c["fixed_amps"], w, "blocks.0.inception.", act)
        assert _rel(tr.out, c["block_fixed_out"]) < 1e-5, name
        assert tr.groups.periods == c["fixed_group_periods"], name
        tf = orc.timesblock_forward(c["x"], w, "blocks.0.inception.", wl.k_periods, wl.T, wl.min_period_threshold, act)
        assert tf.selection.periods.tolist() == c["fft_periods"].tolist(), name
        assert _rel(tf.out, c["block_fft_out"]) < 1e-5, name


@pytest.mark.parametrize("key", ["toy.planted.f32", "toy.white.f32", "toy_bf16.white.bf16", "mid.white.f32",
                                 "etth1.planted.f32", "elec.white.bf16"])
def test_stack(golden_dir, key):
    c = torch.load(golden_dir / "stack.pt")[key]
    wd = dict(c["workload"])
    wd["kernel_set"] = tuple(tuple(k) for k in wd["kernel_set"])
    wl = syn.Workload(**wd)
    w = syn.stack_weights(wl, seed=c["weight_seed"])
    dname = key.split(".")[2]
    x = (syn.planted_features(wl.B, wl.T, wl.d_model, 0) if c["input"] == "planted"
         else syn.white_features(wl.B, wl.T, wl.d_model, 1)).to(syn.torch_dtype(dname))
    trace = []
    y = orc.stack_forward(x, w, wl.n_layers, wl.k_periods, wl.T, wl.min_period_threshold, trace=trace)
    assert [t.selection.periods.tolist() for t in trace] == c["periods"]
    tol = 1e-5 if dname == "f32" else 1.6e-2
    assert _rel(_sub(y), c["out_sub"]) <= tol
    if c["out_full"] is not None:
        assert _rel(y, c["out_full"]) <= tol


def test_model_toy(golden_dir):
    g = torch.load(golden_dir / "model_toy.pt")
    for name in ("toy_direct", "toy_longhist", "toy_recursive"):
        c = g[name]
        wd = dict(c["workload"])
        wd["kernel_set"] = tuple(tuple(k) for k in wd["kernel_set"])
        wl = syn.Workload(**wd)
        cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu",
                           wl.min_period_threshold, 1e-3, wl.context_rank > 0, wl.context_rank)
        x = syn.planted_series(wl.B, c["T_in"], wl.N, seed=c["x_seed"])
        r, d = orc.timesnet_forward(x, c["state"], cfg, series_static=c["static"], series_ids=c["ids"],
                                    min_sigma_vector=c["min_sigma_vector"])
        assert _rel(r, c["rate"]) < 1e-5 and _rel(d, c["disp"]) < 1e-5, name
        assert _rel(orc.nb_nll(c["y"], r, d, c["mask"]), c["nll"]) < 1e-5, name
        if wl.mode == "recursive":
            rr, rd = orc.forecast_recursive(x, wl.H, c["state"], cfg, series_static=c["static"], series_ids=c["ids"])
            assert _rel(rr, c["rec_rate"]) < 2e-5 and _rel(rd, c["rec_disp"]) < 2e-5
    c = g["lowrank"]
    assert _rel(orc.lowrank_context(c["coeff"], c["length"], torch.tensor(c["scale"])), c["ctx"]) < 1e-6
    c = g["nll_small"]
    assert torch.isnan(orc.nb_nll(c["y"], c["rate"], c["disp"], c["mask"]))      # NaN target poisons the masked sum
    assert _rel(orc.nb_nll(c["y_finite"], c["rate"], c["disp"], c["mask"]), c["nll"]) < 1e-6
    assert _rel(orc.nb_nll(c["y_finite"], c["rate"], c["disp"]), c["nll_nomask"]) < 1e-6


# --------------------------------------------------------------------------- #
# round 2 fixtures (oracle/make_golden_r2.py)
# --------------------------------------------------------------------------- #
@pytest.mark.parametrize("key", ["traffic.planted.f32", "traffic.white.bf16"])
def test_r2_traffic_stack(golden_dir, key):
    c = torch.load(golden_dir / "r2_stack_traffic.pt")[key]
    d = dict(c["workload"])
    d["kernel_set"] = tuple(tuple(k) for k in d["kernel_set"])
    wl = syn.Workload(**d)
    dname = key.split(".")[2]
    w = syn.stack_weights(wl, seed=c["weight_seed"])
    x = (syn.planted_features(wl.B, wl.T, wl.d_model, 0) if c["input"] == "planted"
         else syn.white_features(wl.B, wl.T, wl.d_model, 1)).to(syn.torch_dtype(dname))
    trace = []
    y = orc.stack_forward(x, w, wl.n_layers, wl.k_periods, wl.T, wl.min_period_threshold, trace=trace)
    assert [t.selection.periods.tolist() for t in trace] == c["periods"]
    assert _rel(_sub(y), c["out_sub"]) < (1e-5 if dname == "f32" else 1.6e-2)


def test_r2_recursive5_forward(golden_dir):
    c = torch.load(golden_dir / "r2_recursive5.pt")
    d = dict(c["workload"])
    d["kernel_set"] = tuple(tuple(k) for k in d["kernel_set"])
    wl = syn.Workload(**d)
    g = torch.Generator().manual_seed(c["x_seed"])
    x = torch.poisson(torch.full((wl.B, wl.T, wl.N), 4.0), generator=g)
    static = torch.randn(wl.B, wl.N, wl.static_features, generator=g)
    sd = syn.seeded_state(c["state_keys"], seed=c["state_seed"])
    cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu", wl.min_period_threshold, 1e-3,
                       True, wl.context_rank)
    n = 64                                              # period selection is shared: keep the whole batch for it
    r, dd = orc.timesnet_forward(x, sd, cfg, series_static=static, series_ids=torch.arange(wl.N))
    assert _rel(r, c["rate"]) < 1e-5 and _rel(dd, c["disp"]) < 1e-5
    assert _rel(r[:n], c["rec_rate"][:n, :1]) < 1e-5


def test_r2_embedding_with_marks(golden_dir):
    g = torch.load(golden_dir / "r2_embed_mark.pt")
    for mode in ("decoupled", "none", "layer", "rms"):
        c = g[f"embed.{mode}"]
        w = {"embedding." + k: v for k, v in c["state"].items()}
        assert _rel(orc.data_embedding(c["x"], w, c["mark"], mode), c["out"]) < 1e-6, mode
        assert _rel(orc.data_embedding(c["x"], w, None, mode), c["out_nomark"]) < 1e-6, mode
    for name in ("mark_direct", "mark_recursive"):
        c = g[name]
        d = dict(c["workload"])
        d["kernel_set"] = tuple(tuple(k) for k in d["kernel_set"])
        wl = syn.Workload(**d)
        cfg = orc.ModelCfg(wl.T, wl.H, wl.d_model, wl.n_layers, wl.k_periods, wl.mode, "gelu", wl.min_period_threshold,
                           1e-3, wl.context_rank > 0, wl.context_rank)
        x = syn.planted_series(wl.B, c["T_in"], wl.N, seed=c["x_seed"])
        r, dd = orc.timesnet_forward(x, c["state"], cfg, x_mark=c["x_mark"], series_ids=torch.arange(wl.N))
        assert _rel(r, c["rate"]) < 1e-5 and _rel(dd, c["disp"]) < 1e-5, name


def test_r2_block_env_modes(golden_dir):
    g = torch.load(golden_dir / "r2_block_env.pt")
    for name, c in g.items():
        d = dict(c["workload"])
        d["kernel_set"] = tuple(tuple(k) for k in d["kernel_set"])
        wl = syn.Workload(**d)
        w = syn.stack_weights(wl, seed=c["weight_seed"])
        for depth, r in c["by_depth"].items():
            base = 2.0 if "TIMES_PERIOD_BINNING" in c["env"] else None
            mu = None
            if "TIMES_PERIOD_MAX_UNIQ" in c["env"]:
                raw = c["env"]["TIMES_PERIOD_MAX_UNIQ"]
                mu = int(raw) if raw.isdigit() else {0: 3, 1: 1}[depth]
            tr = orc.timesblock_from_periods(c["x"], c["periods"], c["amps"], w, "blocks.0.inception.", "gelu",
                                             log_base=base, max_unique=mu)
            assert len(tr.groups.periods) == r["groups"], (name, depth)
            assert _rel(tr.out, r["fixed_out"]) < 1e-5, (name, depth)


def test_r2_inception_nchw(golden_dir):
    g = torch.load(golden_dir / "r2_inception_nchw.pt")
    for name, c in g.items():
        if c["kind"] == "block":
            y = orc.inception_block(c["x"], c["state"], "", c["act"])
            assert _rel(y, c["out"]) < 1e-5, name
        elif c["kind"] == "rms":
            assert _rel(orc.rms_norm(c["x"], c["state"]["weight"], c["state"]["bias"]), c["out"]) < 1e-6
