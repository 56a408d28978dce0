#!/bin/bash
# Round-2 GPU pass R: gpu suite + default bench + reference arm on the final build (one-row-per-lane line search, obs_policy_ld).
tag=${1:-r02r}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -m gpu -q -s -rxX > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED" $o/${tag}_pytest_gpu.log | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -n 4 $o/${tag}_smoke.log
timeout 600 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; cut -c1-160 $o/${tag}_bench_ref.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step -f python bench.py --pipeline 1 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step.ncu-rep
ls $o | grep ${tag}
