#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1
timeout 200 python tools/bench_mel.py > gpurun_out/bench_mel.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 1 gpurun_out/bench_train.log | cut -c1-700; tail -n 1 gpurun_out/bench_mel.log | cut -c1-700
