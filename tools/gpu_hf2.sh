#!/bin/bash
# height-field instantiation: parity tests, rollout bench on rough_terrain_backlash (4096 and 2048 envs/GPU), launch list; then smoke, all gpu tests and the default bench line
tag=${1:-r01h}
o=gpurun_out
mkdir -p $o
timeout 300 python -m pytest tests/test_hfield.py -m gpu -q -s > $o/${tag}_pytest_hfield.log 2>&1; echo "hfield pytest exit $?" >> $o/${tag}_pytest_hfield.log; tail -3 $o/${tag}_pytest_hfield.log | cut -c1-300
timeout 200 python bench.py --task rough_terrain_backlash --steps 100 --warmup 10 --no-cpu-baseline > $o/${tag}_bench_rough_n1.json 2> $o/${tag}_bench_rough_n1.err; cut -c1-300 $o/${tag}_bench_rough_n1.json; tail -2 $o/${tag}_bench_rough_n1.err
timeout 200 python bench.py --task rough_terrain_backlash --envs-per-gpu 2048 --steps 100 --warmup 10 --no-cpu-baseline > $o/${tag}_bench_rough_2048.json 2> $o/${tag}_bench_rough_2048.err; cut -c1-300 $o/${tag}_bench_rough_2048.json
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/${tag}_launches_rough.csv python bench.py --task rough_terrain_backlash --steps 4 --warmup 3 --no-cpu-baseline > $o/${tag}_launches_rough.log 2>&1; grep -c k_step $o/${tag}_launches_rough.csv
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout 300 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit" $o/${tag}_pytest_gpu.log | tail -3
timeout 200 python bench.py --steps 200 --warmup 20 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json
ls -la $o | grep ${tag}
