python -m pytest tests -m gpu -x -q 2>&1 | tail -8
AMT_TRACE=1 python bench.py --fovs 32 --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep amt-trace | tail -9
python bench.py --fovs 64 --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','fov_per_s','cells_per_fov')})
    else: print(l.rstrip()[:300])
"
