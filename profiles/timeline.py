"""Kernel timeline of the stack step replayed from its CUDA graph (CUPTI activity records through torch.profiler: start,
duration and stream of every kernel with concurrency preserved -- ncu serialises, nsys is not in the image).
usage: python profiles/timeline.py [workload] [e2e]   -> gpurun_out/timeline_<workload>.txt"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
import bench  # noqa: E402
import flowtimes_synth as syn  # noqa: E402
from timesnet_forecast.cuda_graphs import GraphedCallable  # noqa: E402
from timesnet_forecast.models.timesnet import TimesNet  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "elec"
e2e = len(sys.argv) > 2 and sys.argv[2] == "e2e"
wl = syn.WORKLOADS[name]
dev = torch.device("cuda", 0)
sdt = syn.torch_dtype(wl.dtype)
model = TimesNet(input_len=wl.T, pred_len=wl.H, d_model=wl.d_model, n_layers=wl.n_layers, k_periods=wl.k_periods,
                 kernel_set=[list(k) for k in wl.kernel_set], dropout=0.0, activation="gelu", mode=wl.mode, d_ff=wl.ff,
                 bottleneck_ratio=wl.bottleneck_ratio, min_period_threshold=wl.min_period_threshold, use_checkpoint=False,
                 stack_dtype=sdt)
x = syn.planted_series(wl.B, wl.T, wl.N, seed=0).to(dev)
model.eval()
model(x[:1])
model.load_state_dict(bench.model_state(wl, dev), strict=True)
model.check_finite = False
feats = [syn.planted_features(wl.B, wl.T, wl.d_model, seed=i).to(sdt).to(dev) for i in range(4)]
if e2e:
    fn, inputs = (lambda xx: model(xx)), [x]
else:
    fn, inputs = model.stack_forward, [feats[0]]
for _ in range(3):
    fn(*inputs)
g = GraphedCallable(fn, inputs, params_of=model)
for i in range(5):
    g(*inputs)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        g(*inputs)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# last replay only
starts = [e for e in evs if "tc_dft" in e.name or "spectrum_fft" in e.name or "Memcpy" in e.name]
t0 = evs[0].time_range.start
out = Path(ROOT / "gpurun_out" / f"timeline_{name}{'_e2e' if e2e else ''}.txt")
with out.open("w") as f:
    f.write("# start_us  dur_us  end_us  stream  kernel\n")
    for e in evs:
        s = e.time_range.start - t0
        d = e.time_range.end - e.time_range.start
        nm = e.name.replace("void ", "").replace("ftn::", "")[:70]
        f.write(f"{s:10.1f} {d:8.1f} {s + d:10.1f}  {getattr(e, 'device_index', 0)}  {nm}\n")
print(out.read_text()[-6000:])
