#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/smoke.log | cut -c1-200; grep -a "mel (512" gpurun_out/pytest_gpu.log; tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-200
