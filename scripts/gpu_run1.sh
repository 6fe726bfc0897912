set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2a.log 2>&1; echo "pytest rc=$?"
python bench.py --fovs 16 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_s2a_small.json 2> gpurun_out/bench_s2a_small.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dog_pass_kernel|gauss_h_kernel' -s 2 -c 2 -o gpurun_out/prof_dog_s2a -f python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_s2a.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'map_kernel|sel_hist|sel_compact|ccl_tile|ccl_compress|region_reduce|relabel_final|otsu' -s 8 -c 12 -o gpurun_out/prof_rest_s2a -f python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_s2a_rest.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/pytest_gpu_s2a.log; cat gpurun_out/bench_s2a_small.json | head -c 600
