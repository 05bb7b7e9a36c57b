#!/bin/bash
mkdir -p gpurun_out
python tools/conv_microbench.py > gpurun_out/mb_default.log 2>&1; tail -14 gpurun_out/mb_default.log
for st in 4 2; do IISEG_CONV_STAGES=$st python tools/conv_microbench.py > gpurun_out/mb_stages$st.log 2>&1; tail -14 gpurun_out/mb_stages$st.log; done
for d in 1 2 3; do IISEG_CONV_DBG=$d python tools/conv_microbench.py > gpurun_out/mb_dbg$d.log 2>&1; tail -14 gpurun_out/mb_dbg$d.log; done
