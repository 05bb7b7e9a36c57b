#!/bin/bash
# One gpurun call: bring-up + kernel tests + path tests + parity numbers, each in its own process.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run bringup python tools/bringup_conv.py
TAILN=15 run stream python -m pytest tests/test_streaming_kernels_gpu.py -m gpu -q -x --timeout 300
TAILN=30 run conv python -m pytest tests/test_conv_gpu.py -m gpu -q --timeout 300
TAILN=40 run path python -m pytest tests/test_path_gpu.py -m gpu -q --timeout 600
TAILN=30 run parity python tools/parity_report.py 64 80 6
