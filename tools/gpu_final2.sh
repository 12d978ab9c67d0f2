#!/bin/bash
# round-end dry run (tools/gpu_final.sh) + the timing tools and one ncu capture of the train()-mode and log-mel kernels
bash tools/gpu_final.sh > gpurun_out/final.log 2>&1; tail -n 9 gpurun_out/final.log | cut -c1-200
timeout 300 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; tail -n 1 gpurun_out/bench_train.log | cut -c1-600
timeout 300 python tools/bench_mel.py > gpurun_out/bench_mel.log 2>&1; tail -n 1 gpurun_out/bench_mel.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -o gpurun_out/prof_r01h_train_mel -f python tools/prof_new.py > gpurun_out/ncu_train_mel.log 2>&1; tail -n 2 gpurun_out/ncu_train_mel.log | cut -c1-200
ls -la gpurun_out/prof_r01h_train_mel.ncu-rep
