#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1; echo "rc=$?" >> gpurun_out/bench_short.log
tail -n 4 gpurun_out/smoke.log | cut -c1-200; tail -n 3 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/bench_short.log | cut -c1-260
