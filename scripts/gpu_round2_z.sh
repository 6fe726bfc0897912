#!/bin/bash
# round 2, last evidence pass: full GPU test suite, smoke, bench (N=1, driver-style), reference arm, launch list of a
# two-chunk batch, ncu --set full of the main kernels of one chunk
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02f_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -1 gpurun_out/r02f_smoke.log
timeout 900 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; tail -3 gpurun_out/r02f_bench_n1.err; head -c 300 gpurun_out/r02f_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; head -c 300 gpurun_out/r02f_bench_reference.json
timeout 300 python scripts/prof_chunk.py 16 3 > gpurun_out/r02f_prof_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:amt:: -c 2000 --csv --log-file gpurun_out/r02f_launches.csv python scripts/prof_chunk.py 16 2 > gpurun_out/r02f_ncu_launches.log 2>&1; tail -1 gpurun_out/r02f_prof_plain.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'amt::.*(tcg_axis1|tcg_axis0|sel_hist_buckets|sel_compact_buckets|rank_collect|exact_eval|map_kernel|ccl_tile|ccl_seam|ccl_final|region_reduce|relabel_final)' -s 17 -c 17 -o gpurun_out/r02f_chunk -f python scripts/prof_chunk.py 8 2 > gpurun_out/r02f_ncu_chunk.log 2>&1; tail -1 gpurun_out/r02f_ncu_chunk.log; ls -la gpurun_out
