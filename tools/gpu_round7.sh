#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=25 run bringup python tools/bringup_conv.py
TAILN=25 run tests python -m pytest tests -m gpu -q --timeout 900
TAILN=14 run mb python tools/conv_microbench.py
TAILN=3 run bench python bench.py
