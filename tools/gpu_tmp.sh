timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01l_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01l_pytest_gpu.log; tail -4 gpurun_out/r01l_pytest_gpu.log
ODUCK_CUDA_LIB=open_duck_playground_b200/csrc/variants/liboduck_cuda_precdiv.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_standing.py -m gpu -q > gpurun_out/r01l_pytest_gpu_precdiv.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01l_pytest_gpu_precdiv.log; tail -4 gpurun_out/r01l_pytest_gpu_precdiv.log
python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r01l_bench_n1.json 2>gpurun_out/r01l_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01l_bench_n1.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'kstep', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
python tools/variants.py bench --steps 200 --warmup 20 --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/r01l_launches_ppo.csv python bench.py --mode ppo --steps 20 --warmup 1 > gpurun_out/r01l_launches_ppo.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -o gpurun_out/r01l_k_step -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r01l_ncu_full.log 2>&1
