#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=4 run tests python -m pytest tests -m gpu -q -x --timeout 600
TAILN=3 run bench python bench.py
echo "=== ncu"
python tools/one_application.py > gpurun_out/oneapp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/one_application.py > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?"
