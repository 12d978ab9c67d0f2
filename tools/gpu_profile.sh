#!/bin/bash
# Per-op table + ncu launch list + one full ncu capture of the tap-GEMM kernels.
mkdir -p gpurun_out
timeout 300 python tools/op_table.py --config cfg2 --out gpurun_out/ops_cfg2_rb2.json > gpurun_out/ops_cfg2_rb2.txt 2>&1
timeout 300 python tools/op_table.py --config cfg2 --resblock 1 --out gpurun_out/ops_cfg2_rb1.json > gpurun_out/ops_cfg2_rb1.txt 2>&1
timeout 300 python tools/op_table.py --config cfg3 --out gpurun_out/ops_cfg3_rb2.json > gpurun_out/ops_cfg3_rb2.txt 2>&1
timeout 300 python tools/op_table.py --config cfg2 --precision fp32 --out gpurun_out/ops_cfg2_rb2_fp32.json > gpurun_out/ops_cfg2_rb2_fp32.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv \
   --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tapgemm_tc -s 80 -c 36 \
   -o gpurun_out/prof_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/ops_cfg2_rb2.txt gpurun_out/ops_cfg2_rb1.txt gpurun_out/ops_cfg3_rb2.txt gpurun_out/ops_cfg2_rb2_fp32.txt
