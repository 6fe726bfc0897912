#!/bin/bash
# round 2, pass B: full GPU test suite, smoke, bench (N=1), launch list of one bench step, ncu of the DoG kernels
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_b.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu_b.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_b.log 2>&1; tail -2 gpurun_out/r02_smoke_b.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_b_n1.json 2> gpurun_out/r02_bench_b_n1.err; tail -5 gpurun_out/r02_bench_b_n1.err; head -c 1500 gpurun_out/r02_bench_b_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_b_reference.json 2> gpurun_out/r02_bench_b_reference.err; head -c 600 gpurun_out/r02_bench_b_reference.json
timeout 300 python bench.py --steps 1 --warmup 3 --fovs 32 --e2e-fovs 32 --no-cpu --no-modes > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_b.csv python bench.py --steps 1 --warmup 3 --fovs 32 --e2e-fovs 32 --no-cpu --no-modes > gpurun_out/ncu_launches.log 2>&1
timeout 300 python scripts/prof_dog_chunk.py > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tcg_|lo2d|dog_strip" -c 12 -o gpurun_out/r02_dog_chunk_b python scripts/prof_dog_chunk.py > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
