#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mel.py -m gpu -q -x -s > gpurun_out/pytest_mel.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mel.log
timeout 300 python tools/bench_mel.py > gpurun_out/bench_mel.log 2>&1; echo "rc=$?" >> gpurun_out/bench_mel.log
tail -n 30 gpurun_out/pytest_mel.log | cut -c1-220; tail -n 3 gpurun_out/bench_mel.log | cut -c1-900
