#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2.log
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n1.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref_n2.log
cat gpurun_out/gpus.txt; tail -n 3 gpurun_out/bench_n2.log | cut -c1-600; tail -n 2 gpurun_out/bench_n1.log | cut -c1-300; tail -n 2 gpurun_out/bench_ref_n2.log | cut -c1-300
