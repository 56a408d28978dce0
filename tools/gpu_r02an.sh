#!/bin/bash
# Round-2 GPU pass AN (the round's last GPU seconds): the product's own train() loop -- CUDA env + device learner + evaluator --
# for as many env-steps as fit under a 17 s wall clock; the curve file is rewritten at every evaluation.
o=gpurun_out; mkdir -p $o
timeout 17 python tools/train_curve_cpu.py --device cuda:0 --num_envs 8192 --num_minibatches 32 --num_eval_envs 128 --num_timesteps 60000000 --num_evals 14 \
  --learner_matmul tf32 --out $o/r02an_gpu_learning_curve.json > $o/r02an_train.log 2>&1; echo "exit $?" >> $o/r02an_train.log
tail -n 4 $o/r02an_train.log | cut -c1-250
