import sys, os
sys.path.insert(0, "flow-timesnet_b200"); sys.path.insert(0, ".")
import torch, bench, flowtimes_synth as syn
from timesnet_forecast.models.timesnet import TimesNet
from timesnet_forecast import _native as nv
wl = syn.WORKLOADS["elec"]
m = TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods, kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0, activation="gelu", mode=wl.mode, d_ff=wl.ff, bottleneck_ratio=wl.bottleneck_ratio, min_period_threshold=wl.min_period_threshold, use_checkpoint=False, stack_dtype=torch.bfloat16)
x = syn.planted_series(wl.B, wl.T, wl.N, seed=0)
m(x[:1].cuda()); m.eval(); m.load_state_dict(bench.model_state(wl, torch.device("cuda")), strict=True); m.check_finite=False
B = int(os.environ.get("DBG_B", "64"))
for it in range(3):
    r, d = m(x[:B].cuda())
    torch.cuda.synchronize()
    print("iter", it, "ok", float(r.mean()))
for b in m.blocks: print("periods", b.period_selector.last_selected_periods.tolist())
