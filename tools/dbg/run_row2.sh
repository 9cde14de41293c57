timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x --timeout=300 -k "streaming or etth1 or per_period_delta or stack_golden" > gpurun_out/row2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/row2_pytest.log
for dt in f32 bf16; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload etth1 --dtype $dt --no-cpu-baseline --no-e2e > gpurun_out/row2_bench_etth1_$dt.json 2> gpurun_out/row2_bench_etth1_$dt.err; echo "etth1 $dt rc=$?"; tail -2 gpurun_out/row2_bench_etth1_$dt.err
done
python - <<'PY'
import json
for t in ["f32","bf16"]:
    try:
        d=json.load(open(f"gpurun_out/row2_bench_etth1_{t}.json"))
        ch={k["kernel"][:8]:round(k["avg_ms"]*1e3,1) for k in d["roofline"]["chain_kernels"]}
        print(t, round(d["ms_per_step"],4), round(d["value"]), ch)
    except Exception as e: print(t, "failed", e)
PY
