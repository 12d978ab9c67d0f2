#!/bin/bash
# ncu launch list of a bench run + one full capture of the tap-GEMM launches of one forward.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/op_table.py --config cfg2 --out gpurun_out/ops_cfg2_rb2.json > gpurun_out/ops_cfg2_rb2.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 80 --csv \
   --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tapgemm_tc|conv_post|cond_affine|transpose" -s 120 -c 39 \
   -o gpurun_out/prof_v2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --cpu-seconds 5 > gpurun_out/bench.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/ops_cfg2_rb2.txt; ls -la gpurun_out | head -30
