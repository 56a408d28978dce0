timeout 600 python -m pytest tests/test_ppo_device.py -m gpu -q > gpurun_out/r01s_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01s_pytest.log; tail -3 gpurun_out/r01s_pytest.log
python bench.py --mode ppo --steps 100 --warmup 2 2>/dev/null | tee gpurun_out/r01s_bench_ppo.json | grep -o '"value.\{20\}\|"split.*'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 300 --csv --log-file gpurun_out/r01s_launches_ppo.csv python bench.py --mode ppo --steps 20 --warmup 1 > /dev/null 2>&1
