#!/bin/bash
# Round-2 GPU pass AH: env kernels built with --use_fast_math: full gpu suite, smoke, default bench line, rough bench.
tag=${1:-r02ah}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -q -m gpu -s > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit|FAILED" $o/${tag}_pytest_gpu.log | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -n 3 $o/${tag}_smoke.log
timeout 600 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
