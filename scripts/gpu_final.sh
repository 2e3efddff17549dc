mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log | cut -c1-300
timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_default.json
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --row-shards 2 --no-cpu-baseline > gpurun_out/bench_g2_rows.json 2> gpurun_out/bench_g2_rows.err; echo "bench2 rows rc=$?"; grep -v "^NCCL" gpurun_out/bench_g2_rows.json | cut -c1-200
python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" 2>&1 | tail -2
