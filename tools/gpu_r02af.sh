#!/bin/bash
# Round-2 GPU pass AF: lane-parallel triangle enumeration in the HF collider: height-field parity tests + rough bench.
tag=${1:-r02af}
o=gpurun_out
mkdir -p $o
timeout 600 python -m pytest tests/test_hfield.py tests/test_ppo_device.py -m gpu -q 2>&1 | tail -2
for E in 4096 16384; do
  timeout 300 python bench.py --mode rough --rough-envs $E --steps 40 > $o/${tag}_bench_rough_$E.json 2> $o/${tag}_bench_rough_$E.err; python -c "import json,sys; j=json.load(open('$o/${tag}_bench_rough_$E.json')); print('rough', $E, j['value'], j['ms_per_step'])"
done
