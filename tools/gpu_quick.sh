#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/op_table.py --config cfg2 --out gpurun_out/ops_cfg2_rb2.json > gpurun_out/ops_cfg2_rb2.txt 2>&1
timeout 300 python tools/op_table.py --config cfg5_per_gpu --out gpurun_out/ops_cfg5_rb2.json > gpurun_out/ops_cfg5_rb2.txt 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1
grep -E "fused vs|passed|failed|Error" gpurun_out/pytest_gpu.log | tail -n 8
grep -E "FUSED|total|cond|post" gpurun_out/ops_cfg2_rb2.txt gpurun_out/ops_cfg5_rb2.txt
tail -n 1 gpurun_out/bench.log | cut -c1-200
