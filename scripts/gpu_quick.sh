mkdir -p gpurun_out
{
echo "== c2 shard of 8: 1.25Mx300 m30"; timeout 600 python scripts/sweep_scan.py --rows 1250000 --dim 300 --m 30 --queries 23680 --reps 3 24:0:8192 24:32:8192 24:0:16384 24:32:16384 24:0:32768 24:32:32768 20:32:8192 20:32:16384 20:32:32768 16:32:16384 30:0:32768
echo "== c1"; timeout 300 python scripts/sweep_scan.py --rows 1000000 --dim 100 --m 10 --centres 0 --queries 23680 --reps 3 10:0:8192 10:0:16384 10:0:32768 8:32:8192 8:32:16384
} > gpurun_out/r01e_shapes3.log 2>&1
cat gpurun_out/r01e_shapes3.log
