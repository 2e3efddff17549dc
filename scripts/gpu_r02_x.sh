#!/bin/bash
# round 2, call x: whole GPU suite with the tensor scan selected automatically, then the default bench line
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02zw_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02zw_tests.log | cut -c1-800
timeout -s KILL 900 python bench.py > gpurun_out/r02zw_bench.json 2> gpurun_out/r02zw_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02zw_bench.json; tail -3 gpurun_out/r02zw_bench.err
