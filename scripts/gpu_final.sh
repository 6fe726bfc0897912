# End-of-session measurement pass: tests, default bench, reference arm, launch list, full ncu captures.
set -x
T=${TAG:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${T}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_${T}.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${T}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${T}.log
python bench.py > gpurun_out/bench_${T}_n1.json 2> gpurun_out/bench_${T}_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${T}_reference.json 2> gpurun_out/bench_${T}_reference.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-contracted > gpurun_out/ncu_launches_${T}.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dog_strip|sel_hist|sel_compact|map_kernel|ccl_tile|ccl_seam|ccl_final|relabel_final|region_reduce' -s 17 -c 17 -o gpurun_out/prof_all_${T} -f python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-contracted > gpurun_out/ncu_all_${T}.log 2>&1; echo "ncu rc=$?"
