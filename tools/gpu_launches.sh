#!/bin/bash
mkdir -p gpurun_out
python tools/steady_application.py > gpurun_out/steady.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_steady.csv python tools/steady_application.py > gpurun_out/ncu_l.log 2>&1
echo "ncu exit $?"
