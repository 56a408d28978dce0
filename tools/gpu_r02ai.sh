#!/bin/bash
# Round-2 GPU pass AI: ncu --set full of the final flat / height-field k_step (fast-math build), reference arm.
tag=${1:-r02ai}
o=gpurun_out
mkdir -p $o
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $o/${tag}_k_step -f python bench.py --pipeline 1 --steps 6 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_ncu_k_step.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 1 -o $o/${tag}_k_step_hf -f python bench.py --mode rough --rough-envs 4096 --steps 6 > $o/${tag}_ncu_k_step_hf.log 2>&1; bash tools/ncu_export.sh $o/${tag}_k_step_hf.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_step -s 30 -c 4 --csv --log-file $o/${tag}_traffic_2048.csv python bench.py --pipeline 1 --envs-per-gpu 2048 --steps 8 --warmup 3 --no-cpu-baseline --no-extra > $o/${tag}_traffic_2048.log 2>&1
ls $o | grep ${tag}
