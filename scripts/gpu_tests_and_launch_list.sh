python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_latest.csv python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-contracted > gpurun_out/ncu_launches_latest.log 2>&1; echo rc=$?
