for t in "dog_variant=1" "dog_variant=1,exec_swap_prio=0" "dog_variant=0,dog_ctas=2" "dog_variant=0,dog_ctas=2,exec_swap_prio=0" "dog_variant=0" "dog_variant=0,exec_swap_prio=0" "dog_variant=2,dog_ctas=1,exec_swap_prio=0"; do
  echo "== $t"
  AMT_TUNE="$t" python bench.py --fovs 64 --steps 3 --warmup 3 --no-e2e --no-cpu --no-contracted 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','fov_per_s')})
    else: print(l.rstrip()[:300])
"
done
