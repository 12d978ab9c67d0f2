#!/bin/bash
# ncu launch list of three bench steps + one full capture of every kernel of one forward (cfg2: 18 launches per step).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 108 -c 54 --csv \
   --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tapgemm|mrf_stage|mrf_stacked|conv_post|cond_affine|transpose" -s 54 -c 18 \
   -o gpurun_out/prof_r01b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 --cpu-seconds 10 > gpurun_out/bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1
ls -la gpurun_out | head; tail -n 1 gpurun_out/bench.log | cut -c1-300; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
