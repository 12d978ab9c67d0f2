#!/bin/bash
# ncu capture of one train()-mode forward + one log-mel call, summarised ON the box (the report is > 64 MiB),
# then both bench arms with default flags
mkdir -p gpurun_out /tmp/prof
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -o /tmp/prof/train_mel -f python tools/prof_new.py > gpurun_out/ncu_train_mel.log 2>&1; tail -n 2 gpurun_out/ncu_train_mel.log | cut -c1-200
python tools/ncu_summary.py /tmp/prof/train_mel.ncu-rep > gpurun_out/ncu_train_mel.csv 2>> gpurun_out/ncu_train_mel.log; wc -l gpurun_out/ncu_train_mel.csv
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_default.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_default.log
tail -n 2 gpurun_out/bench_default.log | cut -c1-300
