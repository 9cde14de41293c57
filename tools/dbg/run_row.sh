timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x --timeout=300 > gpurun_out/row_pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -4 gpurun_out/row_pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -q -x --timeout=300 > gpurun_out/row_pytest_par.log 2>&1; echo "pytest parity rc=$?"; tail -4 gpurun_out/row_pytest_par.log
for dt in f32 bf16; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 --dtype $dt --no-cpu-baseline --no-e2e > gpurun_out/row_bench_etth1_$dt.json 2> gpurun_out/row_bench_etth1_$dt.err; echo "etth1 $dt rc=$?"; tail -2 gpurun_out/row_bench_etth1_$dt.err
done
python - <<'PY'
import json
for t in ["f32","bf16"]:
    try:
        d=json.load(open(f"gpurun_out/row_bench_etth1_{t}.json"))
        ch={k["kernel"][:8]:round(k["avg_ms"]*1e3,1) for k in d["roofline"]["chain_kernels"]}
        print(t, round(d["ms_per_step"],4), round(d["value"]), ch)
    except Exception as e: print(t, "failed", e)
PY
