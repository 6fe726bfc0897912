#!/bin/bash
# round 2, pass E: full GPU test suite, smoke, bench (N=1) with the fused pass 2 and chunk 16, reference arm
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02e_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; tail -2 gpurun_out/r02e_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; tail -5 gpurun_out/r02e_bench_n1.err; head -c 1200 gpurun_out/r02e_bench_n1.json
for c in 32; do python bench.py --steps 3 --warmup 3 --fovs 128 --no-e2e --no-cpu --no-modes --chunk $c 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk',d['config'].get('chunk_fovs'),'value',d['value'],'ms/step',d['ms_per_step'])"; done
