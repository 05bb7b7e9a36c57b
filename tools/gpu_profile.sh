#!/bin/bash
# ncu evidence for profiles/: launch lists of the bench command and of one steady-state application, and a
# --set full capture of every kernel of the last steady-state application.  Run under gpurun; every command is
# first run WITHOUT ncu and must exit 0.
set -x
mkdir -p gpurun_out
K='regex:conv_|unpool|norm_finalize|softmax|metrics|pack_kernel|deconv16|onehot|maxpool2'
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_plain.log 2>&1 || exit 1
python tools/steady_application.py > gpurun_out/steady_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 120 --csv \
    --log-file gpurun_out/launches_steady.csv python tools/steady_application.py > gpurun_out/ncu_steady.log 2>&1
# full capture: the last (steady-state) application = 12 convs + 6 unpools after 19 + 18 launches
ncu --set full --clock-control none --import-source on -k 'regex:conv_|unpool' --launch-skip 37 -c 18 -f -o gpurun_out/prof_steady_full \
    python tools/steady_application.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_steady_full.ncu-rep --page raw --csv > gpurun_out/prof_steady_full_raw.csv 2>/dev/null
echo "profile done $?"
