timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01k_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01k_pytest_gpu.log; tail -5 gpurun_out/r01k_pytest_gpu.log
python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r01k_bench_n1.json 2>gpurun_out/r01k_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01k_bench_n1.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'kstep', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
python tools/variants.py bench --steps 200 --warmup 20 --no-cpu-baseline
python bench.py --mode ppo --steps 100 --warmup 2 2>/dev/null | tee gpurun_out/r01k_bench_ppo.json
