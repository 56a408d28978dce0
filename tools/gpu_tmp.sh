o=gpurun_out; tag=r01z
timeout 600 python tools/sweep.py > $o/${tag}_sweep_n1.jsonl 2> $o/${tag}_sweep_n1.err; cat $o/${tag}_sweep_n1.jsonl | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_rollout.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $o/${tag}_launches_rollout.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 1 -o $o/${tag}_k_step -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_k_step.log 2>&1
ls -la $o | grep ${tag}_k_step
