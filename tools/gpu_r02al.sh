#!/bin/bash
# Round-2 GPU pass AL (last 2.6 GPU-minutes of the round): the committed build as the driver runs it -- gpu suite, smoke, default bench line
# (now with roofline.issue) -- then compute-sanitizer memcheck of the final fast-math env kernels (flat + height field) if time is left.
tag=${1:-r02al}
o=gpurun_out
mkdir -p $o
timeout 300 python -m pytest tests -x -q -m gpu > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; tail -n 3 $o/${tag}_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -n 3 $o/${tag}_smoke.log
timeout 200 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; cut -c1-200 $o/${tag}_bench_n1.json; tail -3 $o/${tag}_bench_n1.err
SAN_ENVS=24 timeout 60 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_run.py flat hf > $o/${tag}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> $o/${tag}_sanitizer_memcheck.log; tail -n 4 $o/${tag}_sanitizer_memcheck.log
