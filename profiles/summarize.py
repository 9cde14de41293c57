"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py <tag> [launches.csv] [report.ncu-rep]
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


if __name__ == "__main__":
    tag = sys.argv[1]
    here = Path(__file__).resolve().parent
    if len(sys.argv) > 2 and sys.argv[2] not in ("", "-"):
        launches(Path(sys.argv[2]), here / f"{tag}_launches.txt")
    if len(sys.argv) > 3:
        full(Path(sys.argv[3]), here / f"{tag}_ncu_full.csv")
