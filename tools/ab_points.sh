#!/bin/bash
# A/B of the point kernels on one GPU: tools/ab_bench.py with the default library and with every
# kzg_setup_powersoftau_b200/libptau_b200_*.so, each variant first through the ragged / edge-case parity tests.
OUT=gpurun_out; TAG=${1:-abp}; mkdir -p $OUT
{
echo "== default"; timeout 300 python tools/ab_bench.py 20 2>&1 | grep -v "^imad\|^fq_mul\|dbl loop"
for L in kzg_setup_powersoftau_b200/libptau_b200_*.so; do
  [ -f "$L" ] || continue
  echo "== $L"
  PTAU_LIB=$PWD/$L timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "edge_cases or ragged or random_batches or golden or c_oracle_2pow16" 2>&1 | tail -2
  PTAU_LIB=$PWD/$L timeout 300 python tools/ab_bench.py 20 2>&1 | grep -v "^imad\|^fq_mul\|dbl loop"
done


} | tee $OUT/${TAG}_points.log
