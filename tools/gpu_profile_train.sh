#!/bin/bash
# ncu launch list of the DAE training step (config 4) for profiles/.  The plain run must exit 0 first.
set -x
mkdir -p gpurun_out
python tools/train_bench.py > gpurun_out/train_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k 'regex:conv_|unpool|pool|transpose|rmsprop|loss_grad|noise|bias_grad|sum_slabs' --launch-skip 600 -c 330 --csv \
    --log-file gpurun_out/launches_train.csv python tools/train_bench.py > gpurun_out/ncu_train.log 2>&1
echo "profile done $?"
