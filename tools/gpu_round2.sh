#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=5 run smoke python __graft_entry__.py --smoke
TAILN=6 run bench python bench.py
echo "=== ncu"
python tools/one_application.py > gpurun_out/oneapp.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/one_application.py > gpurun_out/ncu1.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 18 -c 12 -o gpurun_out/prof_conv python tools/one_application.py > gpurun_out/ncu2.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu2.log
