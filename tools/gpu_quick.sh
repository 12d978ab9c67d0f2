#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_probe.py > gpurun_out/probe.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/op_table.py --config cfg2 --out gpurun_out/ops_cfg2_rb2.json > gpurun_out/ops_cfg2_rb2.txt 2>&1
V2W_TC_NO_TMA_OUT=1 timeout 300 python tools/op_table.py --config cfg2 > gpurun_out/ops_cfg2_rb2_notma.txt 2>&1
timeout 300 python tools/op_table.py --config cfg5_per_gpu --out gpurun_out/ops_cfg5_rb2.json > gpurun_out/ops_cfg5_rb2.txt 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --resblock 1 --precision fp16 > gpurun_out/bench_rb1.log 2>&1
tail -n 2 gpurun_out/probe.log; grep BAD gpurun_out/probe.log | head -5
tail -n 3 gpurun_out/pytest_gpu.log
paste -d'|' <(cut -c1-16,95-125 gpurun_out/ops_cfg2_rb2.txt) <(cut -c95-125 gpurun_out/ops_cfg2_rb2_notma.txt) | head -24
tail -n 1 gpurun_out/ops_cfg5_rb2.txt
tail -n 1 gpurun_out/bench.log | cut -c1-200; tail -n 1 gpurun_out/bench_rb1.log | cut -c1-200
