#!/bin/bash
# ncu --set full captures of the two tensor-core kernels (one GPU; run only after the plain commands exited 0).
set -x
python scripts/time_linear.py fwd1 > gpurun_out/plain_fwd1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 2 -c 1 -o gpurun_out/prof_gemm_fwd1_r1b -f python scripts/time_linear.py fwd1 > gpurun_out/ncu_gemm_fwd1.log 2>&1
python scripts/time_topk.py 37888 > gpurun_out/plain_topk.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_topk -s 1 -c 1 -o gpurun_out/prof_topk_r1h -f python scripts/time_topk.py 37888 > gpurun_out/ncu_topk_h.log 2>&1
tail -2 gpurun_out/ncu_gemm_fwd1.log gpurun_out/ncu_topk_h.log
