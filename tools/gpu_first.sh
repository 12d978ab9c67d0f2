#!/bin/bash
# First on-GPU pass: kernel probe, parity tests, smoke, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python tools/tc_probe.py > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "simt or fp32 or sees" > gpurun_out/pytest_fp32.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_fp32.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "not (simt or fp32 or sees)" > gpurun_out/pytest_bf16.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_bf16.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --cpu-seconds 5 > gpurun_out/bench.log 2>&1
echo "rc=$?" >> gpurun_out/bench.log
tail -5 gpurun_out/probe.log gpurun_out/pytest_fp32.log gpurun_out/pytest_bf16.log gpurun_out/smoke.log gpurun_out/bench.log
