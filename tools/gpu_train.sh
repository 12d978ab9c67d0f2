#!/bin/bash
# train()-mode native forward: new tests first, then the whole GPU suite, then the timing tool
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "train_mode" -s > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "rc=$?" >> gpurun_out/bench_train.log
tail -n 25 gpurun_out/pytest_train.log | cut -c1-220; tail -n 4 gpurun_out/pytest_gpu.log | cut -c1-220; tail -n 3 gpurun_out/bench_train.log | cut -c1-900
