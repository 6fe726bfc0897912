SWEEP_CONFIGS="0:0:0,0:0:2,0:0:1,0:1:0,0:1:1,0:2:0,0:2:1" python scripts/dog_sweep.py 32 2>&1 | cut -c1-200 | tail -9
for t in "dog_variant=1" "dog_variant=1,dog_ctas=1" "dog_variant=0" "dog_variant=0,dog_ctas=2" "dog_variant=0,dog_ctas=1" "dog_variant=2,dog_ctas=1" "dog_variant=0,dog_ctas=2,exec_swap_prio=1"; do
  echo "== $t"
  AMT_TUNE="$t" python bench.py --fovs 64 --steps 3 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','fov_per_s','kernels_ms')}, d['roofline_fp64']['frac'])
    else: print(l.rstrip()[:300])
"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
