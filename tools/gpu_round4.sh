#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=6 run tests python -m pytest tests -m gpu -q --timeout 600
TAILN=14 run mb python tools/conv_microbench.py
IISEG_CONV_DBG=3 TAILN=14 run mb3 python tools/conv_microbench.py
TAILN=3 run bench python bench.py
