"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py <tag> [launches.csv] [report.ncu-rep] [workload] [launches_per_step]

With a workload name the --set full capture is also condensed into profiles/<tag>_ncu_summary.json, the file
bench.py reads `roofline.traffic` / `roofline.ncu` from (per-kernel time, DRAM bytes, tensor-pipe activity, and the
step's summed DRAM bytes against the compulsory bytes of SURVEY 8d).
"""
import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def launches(path: Path, out: Path):
    lines = [l for l in path.read_text().splitlines() if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"[<(].*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with out.open("w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {path.name}; total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")
        f.write(f"{'kernel':48s} {'launches':>8s} {'total_us':>11s} {'share':>7s} {'avg_us':>10s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:48s} {n:8d} {t:11.1f} {t / tot:7.3f} {t / n:10.1f}\n")


def full(rep: Path, out: Path):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with out.open("w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source: {rep.name}\n")
        f.write(",".join(f"{hdr[i]} [{units[i]}]" for i in idx) + "\n")
        for r in rows[2:]:
            f.write(",".join(r[i][:60].replace(",", ";") for i in idx) + "\n")


def summary_json(rep: Path, out: Path, workload: str, tag: str):
    import json
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "flow-timesnet_b200"))
    import flowtimes_synth as syn
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in KEYS if k in hdr}

    def num(r, k, scale_unit=True):
        if k not in col:
            return None
        try:
            v = float(r[col[k]].replace(",", ""))
        except ValueError:
            return None
        u = units[col[k]].lower()
        if scale_unit:
            for pre, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
                if u.startswith(pre):
                    return v * f
            if u in ("ns", "nsecond"):
                return v / 1e3
            if u in ("ms", "msecond"):
                return v * 1e3
        return v

    per = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"[<(].*", "", r[col["Kernel Name"]]).replace("void ", "").strip()
        d = per.setdefault(name, dict(launches=0, time_us=0.0, dram_bytes=0.0, tensor=0.0, issue=0.0, fma=0.0, xu=0.0))
        d["launches"] += 1
        d["time_us"] += num(r, "gpu__time_duration.sum") or 0.0
        d["dram_bytes"] += (num(r, "dram__bytes_read.sum") or 0.0) + (num(r, "dram__bytes_write.sum") or 0.0)
        d["tensor"] += num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False) or 0.0
        d["issue"] += num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False) or 0.0
        d["fma"] += num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False) or 0.0
        d["xu"] += num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", False) or 0.0
    kernels = []
    for name, d in per.items():
        n = d["launches"]
        kernels.append({"kernel": name, "launches": n, "avg_us": d["time_us"] / n, "dram_bytes_per_launch": d["dram_bytes"] / n,
                        "tensor_active_pct": d["tensor"] / n, "issue_active_pct": d["issue"] / n,
                        "fma_active_pct": d["fma"] / n, "xu_active_pct": d["xu"] / n})
    wl = syn.WORKLOADS[workload]
    e = 2 if wl.dtype == "bf16" else 4
    # compulsory HBM bytes of the captured stack forward (SURVEY 8d, K3 fused with K4): per layer the search reads x
    # once and the block reads x and writes out once
    compulsory = wl.n_layers * 3 * wl.B * wl.T * wl.d_model * e
    total = sum(d["dram_bytes"] for d in per.values())
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out.write_text(json.dumps({
        "workload": workload, "tag": tag, "git": git, "capture": rep.name,
        "kernels": kernels, "dram_bytes_per_launch": {k["kernel"]: k["dram_bytes_per_launch"] for k in kernels},
        "step_dram_bytes": total, "compulsory_bytes": compulsory, "dram_over_compulsory": total / compulsory,
        "note": "one stack forward captured with ncu --set full --clock-control none (cold, serialised)"}, indent=1) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    here = Path(__file__).resolve().parent
    if len(sys.argv) > 2 and sys.argv[2] not in ("", "-"):
        launches(Path(sys.argv[2]), here / f"{tag}_launches.txt")
    if len(sys.argv) > 3:
        full(Path(sys.argv[3]), here / f"{tag}_ncu_full.csv")
    if len(sys.argv) > 4:
        summary_json(Path(sys.argv[3]), here / f"{tag}_ncu_summary.json", sys.argv[4], tag)
