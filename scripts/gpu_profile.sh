(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/pytest.log
CMD="python bench.py --steps 1 --warmup 1 --queries 4736 --no-cpu-baseline --no-recall"
KR='regex:fused_scan|lut_build|unpack_keys|select_pass|gather_lists|assign_exact|adc_keys|offsets_kernel|rm_|update_|count_diff|gather_rows'
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_scan -s 1 -c 2 -o gpurun_out/prof_scan $CMD > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:assign_exact -s 12 -c 1 -o gpurun_out/prof_encode $CMD > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/plain.log | tail -2; tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
