#!/bin/bash
# last pass of round 2: GPU test suite, smoke, default bench with the shipped library
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02g_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; tail -1 gpurun_out/r02g_smoke.log
timeout 900 python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; tail -3 gpurun_out/r02g_bench_n1.err; head -c 300 gpurun_out/r02g_bench_n1.json
