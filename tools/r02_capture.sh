#!/bin/bash
# Round-2 evidence run on ONE B200 (every command first exits 0 without ncu, as the profiling recipe asks):
#   GPU parity tests, the driver's bench command and its reference arm, the ncu launch list of the bench command,
#   ncu --set full of every heavy kernel at 2^20 points and of the consumer kernels.
# usage: tools/r02_capture.sh [tag=r02] [outdir=gpurun_out]
TAG=${1:-r02}
OUT=${2:-gpurun_out}
mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > $OUT/${TAG}_bench_reference_n1.json 2> $OUT/${TAG}_bench_reference_n1.err; echo "ref rc=$?"
# launch list of the bench command (the recipe's gpu__time_duration pass); config5 shortened to 2^22 powers so the list stays short
python bench.py --steps 2 --warmup 3 --no-legs --log2-powers 22 > $OUT/${TAG}_bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-legs --log2-powers 22 > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
# full capture of the heavy kernels, 2^20 points each (second launch of each = warm)
LOGN=20 python tools/prof_kernels.py > $OUT/${TAG}_prof_kernels.log 2>&1 && \
LOGN=20 ncu --set full --clock-control none -k regex:convert_kernel -o $OUT/${TAG}_ncu_full_kernels -f \
  python tools/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/${TAG}_ncu_full_kernels.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_kernels_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_ncu_full_kernels_raw.csv > $OUT/${TAG}_ncu_full_all_kernels.csv
rm -f $OUT/${TAG}_ncu_full_kernels.ncu-rep   # gpurun_out/ travels back only below 64 MiB
python tools/prof_consumer.py > $OUT/${TAG}_prof_consumer.log 2>&1 && \
ncu --set full --clock-control none -k regex:'msm_|kzg_' -o $OUT/${TAG}_ncu_full_consumer -f \
  python tools/prof_consumer.py > $OUT/${TAG}_ncu_full_consumer.log 2>&1; echo "ncu consumer rc=$?"
ncu -i $OUT/${TAG}_ncu_full_consumer.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_consumer_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_ncu_full_consumer_raw.csv > $OUT/${TAG}_ncu_full_msm_kzg.csv
python tools/kzg_check_bench.py 37888 > $OUT/${TAG}_kzg_check_bench.log 2>&1; cat $OUT/${TAG}_kzg_check_bench.log
python tools/msm_bench.py 16 18 20 22 > $OUT/${TAG}_msm_bench.log 2>&1; cat $OUT/${TAG}_msm_bench.log
rm -f $OUT/${TAG}_ncu_full_consumer.ncu-rep
# A/B builds of the pairing kernels, when present (tools/ab_build_kzg.sh)
for L in kzg_setup_powersoftau_b200/libptau_b200_*.so; do
  [ -f "$L" ] || continue
  echo "== $L"; PTAU_LIB=$PWD/$L python tools/kzg_check_bench.py 37888 2>&1 | tail -2
done | tee $OUT/${TAG}_kzg_ab.log
du -sh $OUT
