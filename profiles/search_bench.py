"""Period search alone at a BASELINE shape: tensor-core DFT route against the SIMT FFT route (CUDA events, warm, inputs
rotating over > 2x L2).  usage: python profiles/search_bench.py [elec|etth1] [iters]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "flow-timesnet_b200"))
import flowtimes_synth as syn  # noqa: E402
from timesnet_forecast import _native as nv  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "elec"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
wl = syn.WORKLOADS[name]
B, L, C, k = wl.B, wl.T, wl.d_model, wl.k_periods
nbuf = max(2, int(300e6 // (B * L * C * 2)))
xs = [syn.planted_features(B, L, C, seed=i).to(torch.bfloat16).cuda() for i in range(min(nbuf, 48))]
for route in (True, False):
    for _ in range(5):
        nv.period_search(xs[0], k, L, 1, tensor_dft=route)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(len(xs)):
            nv.period_search(xs[i], k, L, 1, tensor_dft=route)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name} search route={'tensor' if route else 'simt'}: {e0.elapsed_time(e1) * 1e3 / (iters * len(xs)):.2f} us per search "
          f"(graph of {len(xs)} searches)")

if __import__("os").environ.get("FLOWTIMES_DFT_TRACE"):
    import ctypes
    lib = nv.load()
    for rep in range(3):
        nv.period_search(xs[rep], k, L, 1)
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)()
        lib.ftn_debug_dft_trace(buf)
        t = list(buf)
        names = ["start", "ticket", "tail0", "sums", "ranks", "plan", "finish"]
        if __import__("os").environ["FLOWTIMES_DFT_TRACE"] == "2":      # marks 2..6 belong to the second (warm) run
            print(f"first tail ends {(t[7] - t[0]) / 1e3:.2f} us; warm re-run: " +
                  "  ".join(f"{n}={(t[i] - t[7]) / 1e3:.2f}" for i, n in enumerate(names) if i >= 2))
        else:
            print("trace (us from kernel start): " + "  ".join(f"{n}={(t[i] - t[0]) / 1e3:.2f}" for i, n in enumerate(names)))
