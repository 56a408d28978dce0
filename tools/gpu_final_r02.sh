#!/bin/bash
# Round-2 final check of the committed build: what the driver runs at round end (gpu suite, smoke, default bench line, reference arm).
tag=${1:-r02final}
o=gpurun_out
mkdir -p $o
timeout 900 python -m pytest tests -x -q -m gpu > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; tail -n 3 $o/${tag}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -n 3 $o/${tag}_smoke.log
timeout 600 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
timeout 300 python bench.py --impl reference > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; cut -c1-200 $o/${tag}_bench_ref.json
