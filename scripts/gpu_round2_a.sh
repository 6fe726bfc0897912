#!/bin/bash
# round 2, pass A: full GPU test suite, smoke, bench (N=1) with stage profile and parity check, launch list + ncu of the DoG kernels
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_a.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_a.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_a.log 2>&1; tail -2 gpurun_out/r02_smoke_a.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_a_n1.json 2> gpurun_out/r02_bench_a_n1.err; tail -5 gpurun_out/r02_bench_a_n1.err; head -c 3000 gpurun_out/r02_bench_a_n1.json
timeout 300 python scripts/prof_dog_chunk.py > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tcg_|lo2d|dog_strip" -c 12 -o gpurun_out/r02_dog_chunk python scripts/prof_dog_chunk.py > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
