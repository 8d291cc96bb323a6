#!/bin/bash
# smoke(), consumer-side GPU tests, point-kernel A/B (default library against libptau_b200_prev.so), MSM and KZG10 check
# throughput, and one ncu --set full capture of kzg_check_kernel at one full wave.
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r02_smoke.log
python -m pytest tests -x -q -m gpu -k "kzg or pairing or msm or prepare or consumer or config3 or random_batches or edge" > $OUT/abf_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/abf_pytest.log
{
echo "== default"; timeout 300 python tools/ab_bench.py 20 2>&1 | grep -v "^imad\|^fq_mul\|dbl loop"
python tools/msm_bench.py 16 18 20 22 2>&1 | tail -4; python tools/kzg_check_bench.py 37888 2>&1 | tail -1
echo "== prev"; PTAU_LIB=$PWD/kzg_setup_powersoftau_b200/libptau_b200_prev.so timeout 300 python tools/ab_bench.py 20 2>&1 | grep -v "^imad\|^fq_mul\|dbl loop"
PTAU_LIB=$PWD/kzg_setup_powersoftau_b200/libptau_b200_prev.so python tools/msm_bench.py 20 2>&1 | tail -1
PTAU_LIB=$PWD/kzg_setup_powersoftau_b200/libptau_b200_prev.so python tools/kzg_check_bench.py 37888 2>&1 | tail -1
} | tee $OUT/abf_ab.log
bash tools/kzg_capture.sh > $OUT/abf_kzg_capture.log 2>&1; tail -28 $OUT/abf_kzg_capture.log | cut -c1-110
