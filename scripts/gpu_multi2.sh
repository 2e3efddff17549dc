mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log | cut -c1-300
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench2 rc=$?"; cut -c1-700 gpurun_out/bench_g2.json; tail -3 gpurun_out/bench_g2.err
