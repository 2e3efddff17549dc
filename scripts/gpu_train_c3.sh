mkdir -p gpurun_out
{
echo "== c3 training 10Mx300 m30 25 iters, sum mode (shardable, fixed point)"; timeout 600 python scripts/bench_train.py 10000000 300 30 25 1
echo "== c3 training, literal running mean"; timeout 900 python scripts/bench_train.py 10000000 300 30 25 0
echo "== encode c2 10Mx300 m30"; timeout 300 python scripts/bench_encode.py 10000000 300 30 1
echo "== encode c5 1Mx1000 m100"; timeout 300 python scripts/bench_encode.py 1000000 1000 100 1
echo "== encode c4 shard 12.5Mx128 m16"; timeout 300 python scripts/bench_encode.py 12500000 128 16 1
} > gpurun_out/r01e_train_encode.log 2>&1
cat gpurun_out/r01e_train_encode.log
