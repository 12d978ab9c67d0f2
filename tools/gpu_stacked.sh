#!/bin/bash
# bring-up of the stacked fused kernel: parity first, then op tables for schedule variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_stacked.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stacked.log
tail -n 4 gpurun_out/pytest_stacked.log
for v in ${VARIANTS:-default lag1 nostack}; do
  unset V2W_FUSED_SPLIT_SUM V2W_FUSED_NO_STACK V2W_FUSED_MMA_WARPS
  case $v in
    mma1) export V2W_FUSED_MMA_WARPS=1;;
    split0) export V2W_FUSED_SPLIT_SUM=0;;
    split1) export V2W_FUSED_SPLIT_SUM=1;;
    nostack) export V2W_FUSED_NO_STACK=1;;
  esac
  for cfg in cfg2 cfg5_per_gpu; do
    timeout 300 python tools/op_table.py --config $cfg --out gpurun_out/ops_${cfg}_${v}.json > gpurun_out/ops_${cfg}_${v}.txt 2>&1
    echo "== $v $cfg"; grep "FUSED\|total" gpurun_out/ops_${cfg}_${v}.txt | cut -c1-120
  done
done
