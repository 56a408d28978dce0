#!/bin/bash
# GPU pass for the height-field instantiations + full verification: hfield parity tests (verbose), smoke, all gpu tests, headline bench.
tag=${1:-r01h}
o=gpurun_out
mkdir -p $o
timeout 300 python -m pytest tests/test_hfield.py -m gpu -q -s > $o/${tag}_pytest_hfield.log 2>&1; echo "hfield pytest exit $?" >> $o/${tag}_pytest_hfield.log; tail -40 $o/${tag}_pytest_hfield.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout 600 python -m pytest tests -m gpu -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log; grep -E "passed|failed|pytest exit" $o/${tag}_pytest_gpu.log | tail -3
timeout 300 python bench.py --steps 200 --warmup 20 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; python - <<PY
import json
try:
    d = json.load(open("$o/${tag}_bench_n1.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "kstep", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], d["cpu_baseline"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
