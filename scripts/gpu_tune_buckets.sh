python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for t in "exec_buckets=1" "exec_buckets=0"; do echo "== $t"; AMT_TUNE="$t" python bench.py --fovs 64 --steps 3 --warmup 3 --no-e2e --no-cpu --no-contracted 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','fov_per_s','kernels_ms')})
    else: print(l.rstrip()[:300])
"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_s2p.csv python bench.py --fovs 8 --steps 1 --warmup 1 --no-e2e --no-cpu --no-contracted > gpurun_out/ncu_launches_s2p.log 2>&1; echo rc=$?
