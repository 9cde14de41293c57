timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_tc.py -m gpu -q -x --timeout=300 > gpurun_out/gran_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/gran_pytest.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/gran_bench_elec.json 2> gpurun_out/gran_bench_elec.err; echo "elec rc=$?"; tail -3 gpurun_out/gran_bench_elec.err
FLOWTIMES_TILE_MAJOR=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/tile_bench_elec.json 2> gpurun_out/tile_bench_elec.err; echo "elec tile-major rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --workload recursive --no-cpu-baseline --no-e2e > gpurun_out/gran_bench_recursive.json 2> gpurun_out/gran_bench_recursive.err; echo "recursive rc=$?"; tail -3 gpurun_out/gran_bench_recursive.err
python - <<'PY'
import json
for t in ["gran_bench_elec","tile_bench_elec","gran_bench_recursive"]:
    try:
        d=json.load(open(f"gpurun_out/{t}.json"))
        ch={k["kernel"][:8]:round(k["avg_ms"]*1e3,1) for k in d["roofline"]["chain_kernels"]}
        print(t, round(d["ms_per_step"],4), d["value"], d.get("e2e",{}).get("value"), ch)
    except Exception as e:
        print(t, "failed", e)
PY
