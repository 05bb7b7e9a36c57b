#!/bin/bash
mkdir -p gpurun_out
python tools/steady_application.py > gpurun_out/steady.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 12 -c 12 -o gpurun_out/prof_steady python tools/steady_application.py > gpurun_out/ncu_steady.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_steady.log
