#!/bin/bash
# Last GPU call of round 2 (one B200, ~6 minutes of box time left): A/B of the G2 ladder variants, then the GPU suite, the
# driver's bench command and an ncu --set full capture of the two G2 ladder kernels -- all three with the fastest library.
# Every step has its own timeout so that the call ends by itself.
OUT=gpurun_out; TAG=r02c; mkdir -p $OUT
PKG=kzg_setup_powersoftau_b200
timeout 120 python tools/ab_g2.py > $OUT/${TAG}_ab_default.log 2>&1; cat $OUT/${TAG}_ab_default.log
for L in $PKG/libptau_b200_*.so; do
  [ -f "$L" ] || continue
  V=$(basename $L .so); V=${V#libptau_b200_}
  PTAU_LIB=$PWD/$L timeout 60 python tools/ab_g2.py > $OUT/${TAG}_ab_$V.log 2>&1; cat $OUT/${TAG}_ab_$V.log
done
BEST=$(python - <<'P'
import glob, re, os
best, bt = "default", None
res = {}
for f in sorted(glob.glob("gpurun_out/r02c_ab_*.log")):
    v = os.path.basename(f)[len("r02c_ab_"):-4]
    m = re.search(r"g2_comp_strict\s+([0-9.]+) ms", open(f).read())   # the two kernels of the headline step
    u = re.search(r"g1_comp_strict\s+([0-9.]+) ms", open(f).read())
    if m and u:
        res[v] = float(m.group(1)) + float(u.group(1))
if "default" in res:
    bt = res["default"]
    for v, t in res.items():
        if t < bt * 0.995:   # a variant has to win by more than the run-to-run noise
            best, bt = v, t
print(best)
P
)
echo "best=$BEST" | tee $OUT/${TAG}_best.txt
if [ "$BEST" != "default" ]; then export PTAU_LIB=$PWD/$PKG/libptau_b200_$BEST.so; fi
timeout 150 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest_gpu.log
timeout 170 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
try:
    l=json.load(open('gpurun_out/r02c_bench_n1.json'))
    print('value %.3fM e2e %.3fM frac %.3f g1 %.2fM g2 %.2fM'%(l['value']/1e6,l['e2e']['value']/1e6,l['roofline']['frac'],l['per_group']['g1_points_per_s']/1e6,l['per_group']['g2_points_per_s']/1e6))
except Exception as e:
    print("no bench line:", e)
P
ONLY_G2=1 LOGN=20 timeout 100 ncu --set full --clock-control none -k regex:convert_kernel -o $OUT/${TAG}_ncu_full_g2 -f \
  python tools/prof_kernels.py > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 60 ncu -i $OUT/${TAG}_ncu_full_g2.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_g2_raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/${TAG}_ncu_full_g2_raw.csv > $OUT/${TAG}_ncu_full_g2.csv 2>/dev/null
rm -f $OUT/${TAG}_ncu_full_g2.ncu-rep
echo "elapsed ${SECONDS}s"
