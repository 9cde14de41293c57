# conv4 development loop: parity of the k x k variants, bench, then one traced run (trace run is not a timing)
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k conv 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/c4_bench.json').read()); print(d['value'], d['ms_per_step'], [(k['kernel'][:6], round(k['avg_ms']*1e3,1)) for k in d['roofline']['chain_kernels']])"
FLOWTIMES_CONV_TRACE=gpurun_out/c4_trace.txt timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > /dev/null 2> gpurun_out/c4_trace.err; echo "trace rc=$?"
