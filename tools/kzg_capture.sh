mkdir -p gpurun_out

python tools/prof_consumer.py > gpurun_out/r02_prof_consumer2.log 2>&1 && \
ncu --set full --clock-control none -k regex:'kzg_check' -o gpurun_out/r02_ncu_kzg_check_fullwave -f python tools/prof_consumer.py > gpurun_out/r02_ncu_kzg2.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_ncu_kzg_check_fullwave.ncu-rep --page raw --csv > gpurun_out/r02_ncu_kzg_check_fullwave_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_ncu_kzg_check_fullwave_raw.csv > gpurun_out/r02_ncu_full_kzg_check.csv
rm -f gpurun_out/r02_ncu_kzg_check_fullwave.ncu-rep
cat gpurun_out/r02_ncu_full_kzg_check.csv | cut -c1-120
