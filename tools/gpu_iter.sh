#!/bin/bash
# Iteration pass: probe + parity + op tables + bench.
mkdir -p gpurun_out
timeout 600 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/op_table.py --config cfg2 --out gpurun_out/ops_cfg2_rb2.json > gpurun_out/ops_cfg2_rb2.txt 2>&1
timeout 300 python tools/op_table.py --config cfg2 --resblock 1 --out gpurun_out/ops_cfg2_rb1.json > gpurun_out/ops_cfg2_rb1.txt 2>&1
timeout 300 python tools/op_table.py --config cfg3 --out gpurun_out/ops_cfg3_rb2.json > gpurun_out/ops_cfg3_rb2.txt 2>&1
timeout 300 python tools/op_table.py --config cfg4 --out gpurun_out/ops_cfg4_rb2.json > gpurun_out/ops_cfg4_rb2.txt 2>&1
timeout 300 python tools/op_table.py --config cfg5_per_gpu --out gpurun_out/ops_cfg5_rb2.json > gpurun_out/ops_cfg5_rb2.txt 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --cpu-seconds 5 > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
grep -c OK gpurun_out/probe.log; grep BAD gpurun_out/probe.log; tail -n 4 gpurun_out/pytest_gpu.log
for f in gpurun_out/ops_*.txt; do tail -n 1 $f; done
tail -n 2 gpurun_out/bench.log | cut -c1-400
