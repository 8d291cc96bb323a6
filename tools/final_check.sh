#!/bin/bash
# last run of the round on one GPU: the GPU suite, smoke(), the driver's bench command and its reference arm
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 10 --warmup 3 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > $OUT/r02_bench_reference_n1.json 2> $OUT/r02_bench_reference_n1.err; echo "ref rc=$?"
python - <<'P'
import json
l=json.load(open('gpurun_out/r02_bench_n1.json'))
print('value %.3fM e2e %.3fM frac %.3f g1 %.2fM g2 %.2fM msm %.1fM kzg %.0f'%(l['value']/1e6,l['e2e']['value']/1e6,l['roofline']['frac'],l['per_group']['g1_points_per_s']/1e6,l['per_group']['g2_points_per_s']/1e6,l['extra']['kzg10_commit_2^20(msm)']['points_per_s_kernels']/1e6,l['extra']['kzg10_check_37888(pairings, kernel only)']['openings_per_s']))
P
