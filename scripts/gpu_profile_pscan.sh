CMD="python bench.py --steps 1 --warmup 1 --queries 4736 --no-cpu-baseline --no-recall"
$CMD > gpurun_out/plain_p.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pruned_scan -s 1 -c 1 -o gpurun_out/prof_pscan $CMD > gpurun_out/ncu_p.log 2>&1
tail -2 gpurun_out/plain_p.log | cut -c1-600; tail -4 gpurun_out/ncu_p.log
