#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command and of one steady-state application, and
# --set full captures of the dominant kernels.  Run under gpurun AFTER the plain commands have exited 0.
set -x
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 || exit 1
python tools/steady_application.py > gpurun_out/steady_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches_steady.csv python tools/steady_application.py > gpurun_out/ncu_steady.log 2>&1
# full captures: the last (steady-state) application = launches 41..60 of the 3 x 20
ncu --set full --clock-control none --import-source on --launch-skip 40 -c 20 -f -o gpurun_out/prof_steady_full \
    python tools/steady_application.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_steady_full.ncu-rep --page raw --csv > gpurun_out/prof_steady_full_raw.csv 2>/dev/null
echo "profile done $?"
