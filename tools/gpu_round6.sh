#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=25 run tests python -m pytest tests -m gpu -q --timeout 900
TAILN=3 run bench python bench.py
TAILN=4 run parity python tools/parity_report.py 64 80 4
