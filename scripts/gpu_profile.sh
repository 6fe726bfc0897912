# Launch list of one bench step + full captures of the DoG kernels (run only after the plain bench exits 0)
set -x
python bench.py --fovs 16 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_profiled_cmd.json 2> gpurun_out/bench_profiled_cmd.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dog_strip_kernel' -s 2 -c 2 -o gpurun_out/prof_dog_${TAG} -f python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_dog_${TAG}.log 2>&1
