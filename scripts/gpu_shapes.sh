# the other BASELINE shapes: full bound single stage (old behaviour) against the automatic subset/stage mode
mkdir -p gpurun_out
{
echo "== c1 1Mx100 m10 gaussian"; timeout 300 python scripts/sweep_scan.py --rows 1000000 --dim 100 --m 10 --centres 0 --queries 9472 --opt pruned_min_rows=0 10:0 10:0:16384 10:0:8192 0:32:16384 0:8:16384 5:8:16384
echo "== c5 1Mx1000 m100"; timeout 400 python scripts/sweep_scan.py --rows 1000000 --dim 1000 --m 100 --queries 9472 --opt pruned_min_rows=0 100:0 100:0:16384 0:32:16384 0:8:16384 40:8:16384 24:8:16384 24:8:8192
} > gpurun_out/r01d_shapes2.log 2>&1
cat gpurun_out/r01d_shapes2.log
