timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_standing.py -m gpu -q > gpurun_out/r01r_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01r_pytest_gpu.log; tail -4 gpurun_out/r01r_pytest_gpu.log
python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r01r_bench_n1.json 2>gpurun_out/r01r_bench_n1.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r01r_bench_n1.json') if l.startswith('{')][-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'kstep', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
python tools/variants.py bench --steps 200 --warmup 20 --no-cpu-baseline
