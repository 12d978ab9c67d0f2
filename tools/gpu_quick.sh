#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --resblock 1 > gpurun_out/bench_rb1.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --resblock 1 --precision fp16 > gpurun_out/bench_rb1_fp16.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --precision fp16 > gpurun_out/bench_rb2_fp16.log 2>&1
grep -E "SNR|passed|failed" gpurun_out/pytest_gpu.log | tail -n 30
for f in gpurun_out/bench_rb1.log gpurun_out/bench_rb1_fp16.log gpurun_out/bench_rb2_fp16.log; do tail -n 1 $f | cut -c1-160; done
