#!/bin/bash
# Round-2 GPU pass AM (the round's last GPU seconds): smoke() with the PPO leg at unroll 8.
o=gpurun_out; mkdir -p $o
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $o/r02am_smoke.log 2>&1; echo "smoke exit $?" >> $o/r02am_smoke.log; tail -n 4 $o/r02am_smoke.log
