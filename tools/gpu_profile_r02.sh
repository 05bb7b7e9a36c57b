#!/bin/bash
# Round-2 ncu evidence for profiles/: launch lists of the bench command (parity-grade headline) and of one steady-state
# application in 'mixed' precision, and a --set full capture of every kernel of the last steady-state application.
# Run under gpurun; every command is first run WITHOUT ncu and must exit 0.
set -x
export PATH=/usr/local/cuda/bin:$PATH
mkdir -p gpurun_out
K='regex:conv_|unpool|norm_finalize|softmax|metrics|pack_kernel|deconv16|onehot|maxpool2'
python bench.py --steps 2 --warmup 3 --sections headline --no-cpu-baseline > gpurun_out/r02_bench_plain.log 2>&1 || exit 1
python tools/steady_application.py mixed > gpurun_out/r02_steady_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench_mixed.csv python bench.py --steps 2 --warmup 3 --sections headline --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 120 --csv \
    --log-file gpurun_out/r02_launches_steady_mixed.csv python tools/steady_application.py mixed > gpurun_out/r02_ncu_steady.log 2>&1
# full capture: the last (steady-state) application = 12 convs + 6 unpools after 19 + 18 launches
ncu --set full --clock-control none --import-source on -k 'regex:conv_|unpool' --launch-skip 37 -c 18 -f -o gpurun_out/r02_prof_steady_full_mixed \
    python tools/steady_application.py mixed > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_prof_steady_full_mixed.ncu-rep --page raw --csv > gpurun_out/r02_prof_steady_full_mixed_raw.csv 2>/dev/null
# the bf16 variant's logits conv (the N-packed kernel is new this round)
ncu --set full --clock-control none --import-source on -k 'regex:conv_npack' --launch-skip 2 -c 1 -f -o gpurun_out/r02_prof_npack_bf16 \
    python tools/steady_application.py bf16 > gpurun_out/r02_ncu_npack.log 2>&1
ncu -i gpurun_out/r02_prof_npack_bf16.ncu-rep --page raw --csv > gpurun_out/r02_prof_npack_bf16_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
echo "profile done $?"
