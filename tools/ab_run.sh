#!/bin/bash
# A/B on one GPU: the consumer-side GPU tests with the default library, then KZG10 check / MSM throughput with the default
# library and with every kzg_setup_powersoftau_b200/libptau_b200_*.so (tools/ab_build*.sh).  usage: tools/ab_run.sh [tag]
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -x -q -m gpu -k "kzg or pairing or msm or prepare or consumer" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
{
echo "== default"; python tools/kzg_check_bench.py 37888 2>&1 | tail -2; python tools/msm_bench.py 18 20 22 2>&1 | tail -3
for L in kzg_setup_powersoftau_b200/libptau_b200_*.so; do
  [ -f "$L" ] || continue
  echo "== $L"; PTAU_LIB=$PWD/$L python tools/kzg_check_bench.py 37888 2>&1 | tail -2
done
} | tee $OUT/${TAG}_ab.log
