#!/bin/bash
# what the driver runs at round end, on one box: smoke, GPU tests, both bench arms (default flags)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_default.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_default.log
tail -n 2 gpurun_out/smoke.log; tail -n 3 gpurun_out/pytest_gpu.log
tail -n 2 gpurun_out/bench_ref_default.log | cut -c1-400; tail -n 2 gpurun_out/bench_default.log | cut -c1-2500
