# automatic mode on the BASELINE shapes against pinned settings
mkdir -p gpurun_out
{
echo "== c2 10Mx300 m30"; timeout 400 python scripts/sweep_scan.py --reps 3 0:32 16:32 0:32
echo "== c1 1Mx100 m10 gaussian"; timeout 300 python scripts/sweep_scan.py --rows 1000000 --dim 100 --m 10 --centres 0 --queries 23680 10:0 0:32
echo "== c5 1Mx1000 m100"; timeout 400 python scripts/sweep_scan.py --rows 1000000 --dim 1000 --m 100 --queries 9472 100:0 70:0 0:32
echo "== c4 shard 12.5Mx128 m16"; timeout 400 python scripts/sweep_scan.py --rows 12500000 --dim 128 --m 16 --queries 9472 16:0 0:32
echo "== c2 shard of 8: 1.25Mx300 m30"; timeout 400 python scripts/sweep_scan.py --rows 1250000 --dim 300 --m 30 --queries 23680 30:0 20:0 16:0 0:32
} > gpurun_out/r01e_shapes.log 2>&1
cat gpurun_out/r01e_shapes.log
