"""Kernel launch list of ONE TimesNet.forward + NLL at a bench workload (run under ncu --metrics gpu__time_duration.sum)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
import torch
import bench
import flowtimes_synth as syn
from timesnet_forecast.losses import negative_binomial_nll
from timesnet_forecast.models.timesnet import TimesNet
name = sys.argv[1] if len(sys.argv) > 1 else "elec"
wl = syn.WORKLOADS[name]
sdt = syn.torch_dtype(wl.dtype)
m = TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods,
             kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0, activation="gelu", mode=wl.mode, d_ff=wl.ff,
             bottleneck_ratio=wl.bottleneck_ratio, use_checkpoint=False, stack_dtype=sdt).eval()
x = syn.planted_series(wl.B, wl.T, wl.N, seed=0).cuda()
y = syn.poisson_targets(wl.B, wl.H, wl.N, 5.0, seed=2).cuda()
m(x[:1])
m.load_state_dict(bench.model_state(wl, "cuda"), strict=True)
m.check_finite = False
for _ in range(3):
    r, d = m(x)
    negative_binomial_nll(y, r, d)
torch.cuda.synchronize()
torch.cuda.profiler.start()
r, d = m(x)
loss = negative_binomial_nll(y, r, d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
