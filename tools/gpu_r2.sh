#!/bin/bash
# Round-2 GPU session script.  usage: tools/gpu_r2.sh <stage> [...]; stages: test bench sizes ncu multi
mkdir -p gpurun_out /tmp/prof
for stage in "$@"; do
case $stage in
test)
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 25 > gpurun_out/r2_pytest.log; tail -n 6 gpurun_out/r2_pytest.log ;;
testall)
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 60 > gpurun_out/r2_pytest.log; tail -n 12 gpurun_out/r2_pytest.log ;;
bench)
  timeout 600 python bench.py > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2_bench_cfg2.json
  timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2_bench_cfg2_reference.json 2>&1; cut -c1-300 gpurun_out/r2_bench_cfg2_reference.json ;;
sizes)
  for c in cfg1 cfg3 cfg4 cfg5_per_gpu; do
    timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; echo "$c rc=$?"; cut -c1-200 gpurun_out/r2_bench_$c.json
  done
  for c in cfg1 cfg4; do
    timeout 600 python bench.py --config $c --graph --steps 30 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/r2_bench_${c}_graph.json 2> gpurun_out/r2_bench_${c}_graph.err; echo "$c graph rc=$?"; cut -c1-200 gpurun_out/r2_bench_${c}_graph.json
  done
  for c in cfg2 cfg5_per_gpu; do python tools/op_table.py --config $c --out gpurun_out/r2_ops_$c.json > gpurun_out/r2_ops_$c.txt 2>&1; done; tail -n 3 gpurun_out/r2_ops_cfg2.txt ;;
ncu)
  for c in ${NCU_CFGS:-cfg2 cfg3 cfg5_per_gpu}; do
    timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
       -o /tmp/prof/$c -f python tools/prof_forward.py $c > gpurun_out/r2_ncu_$c.log 2>&1; tail -n 1 gpurun_out/r2_ncu_$c.log | cut -c1-200
    python tools/ncu_summary.py /tmp/prof/$c.ncu-rep > gpurun_out/r2_ncu_full_$c.csv 2>> gpurun_out/r2_ncu_$c.log; wc -l gpurun_out/r2_ncu_full_$c.csv
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
       --log-file gpurun_out/r2_launches_$c.csv python tools/prof_forward.py $c > /dev/null 2>&1
  done
  cp /tmp/prof/cfg2.ncu-rep gpurun_out/r2_cfg2.ncu-rep 2>/dev/null; ls -la gpurun_out/*.ncu-rep ;;
one)
  timeout 900 python -m pytest tests -m gpu -x -q -s -k "$ONE" 2>&1 | tail -n 40 > gpurun_out/r2_pytest_one.log; tail -n 30 gpurun_out/r2_pytest_one.log ;;
fp32tc)
  timeout 1200 python -m pytest tests -m gpu -q -s -k "fp32tc or device_fold or test_fp32_plan_max_abs" 2>&1 | grep -v "^$" | tail -n 70 > gpurun_out/r2_pytest_fp32tc.log; tail -n 45 gpurun_out/r2_pytest_fp32tc.log
  for pr in fp32tc fp32; do
    timeout 600 python bench.py --precision $pr --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg2_$pr.json 2> gpurun_out/r2_bench_cfg2_$pr.err; echo "$pr rc=$?"; cut -c1-250 gpurun_out/r2_bench_cfg2_$pr.json
  done
  python tools/op_table.py --config cfg2 --precision fp32tc --out gpurun_out/r2_ops_cfg2_fp32tc.json > gpurun_out/r2_ops_cfg2_fp32tc.txt 2>&1; tail -n 45 gpurun_out/r2_ops_cfg2_fp32tc.txt ;;
multi)
  N=${GPUS:-2}
  timeout 900 python -m pytest tests/test_sharding.py -m gpu -x -q 2>&1 | tail -n 8 > gpurun_out/r2_pytest_multi.log; tail -n 3 gpurun_out/r2_pytest_multi.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?"; cut -c1-600 gpurun_out/r2_bench_n$N.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
     bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/r2_bench_n${N}_reference.json 2>&1; cut -c1-300 gpurun_out/r2_bench_n${N}_reference.json ;;
esac
done
