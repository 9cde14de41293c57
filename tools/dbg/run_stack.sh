timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py tests/test_gpu_tc.py -m gpu -q -x --timeout=300 > gpurun_out/stack_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/stack_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --workload recursive --no-cpu-baseline --no-e2e > gpurun_out/stack_bench_recursive.json 2> gpurun_out/stack_bench_recursive.err; echo "recursive rc=$?"; tail -3 gpurun_out/stack_bench_recursive.err
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/stack_bench_elec.json 2> gpurun_out/stack_bench_elec.err; echo "elec rc=$?"; tail -3 gpurun_out/stack_bench_elec.err
python - <<'PY'
import json
for t in ["stack_bench_elec","stack_bench_recursive"]:
    try:
        d=json.load(open(f"gpurun_out/{t}.json"))
        ch={k["kernel"][:8]:round(k["avg_ms"]*1e3,1) for k in d["roofline"]["chain_kernels"]}
        print(t, round(d["ms_per_step"],4), d["value"], ch, d["roofline"]["frac"])
    except Exception as e:
        print(t, "failed", e)
PY
