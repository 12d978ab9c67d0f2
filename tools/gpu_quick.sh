#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
for cfg in cfg1 cfg2 cfg3 cfg4 cfg5_per_gpu; do
timeout 300 python tools/op_table.py --config $cfg --out gpurun_out/ops_${cfg}_rb2.json > gpurun_out/ops_${cfg}_rb2.txt 2>&1
tail -n 1 gpurun_out/ops_${cfg}_rb2.txt
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; cut -c1-16,60-175 gpurun_out/ops_cfg2_rb2.txt
tail -n 1 gpurun_out/bench.log | cut -c1-200
