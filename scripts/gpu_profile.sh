# Full ncu capture (--set full, source) of exactly one 8-FOV chunk of the executor: the 34 launches of the
# timed step (the first 34 matching launches are the warm-up step).  Run after the plain bench exited 0.
set -x
T=${TAG:-chunk}
python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-contracted > gpurun_out/bench_profiled_cmd.json 2> gpurun_out/bench_profiled_cmd.err || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'dog_strip|minmax_init|sel_|plan_dog|map_kernel|otsu_kernel|ccl_|acc_init|region_|scan_kernel|relabel_final' \
  -s 34 -c 34 -o gpurun_out/prof_chunk_${T} -f python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-contracted > gpurun_out/ncu_chunk_${T}.log 2>&1
echo "ncu rc=$?"
