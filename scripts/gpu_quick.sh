mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_grouped.py -x -q -k "query or grouped" > gpurun_out/r01e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r01e_pytest.log
{
echo "== c2 shard of 8: 1.25Mx300 m30"; timeout 400 python scripts/sweep_scan.py --rows 1250000 --dim 300 --m 30 --queries 23680 30:0 20:0 0:32 0:32
echo "== c2 10Mx300 m30"; timeout 400 python scripts/sweep_scan.py --reps 3 16:32 0:32 0:32
} > gpurun_out/r01e_shapes2.log 2>&1
cat gpurun_out/r01e_shapes2.log
