for t in "pass_ctas=8" "pass_ctas=2" "pass_ctas=1" "dog_variant=0,pass_ctas=2" "dog_variant=0,pass_ctas=1" "dog_variant=0,dog_ctas=2,pass_ctas=2" "dog_variant=0,dog_ctas=2,pass_ctas=4"; do
  echo "== $t"
  AMT_TUNE="$t" python bench.py --fovs 64 --steps 3 --warmup 3 --no-e2e --no-cpu --no-contracted 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print({k: d[k] for k in ('value','ms_per_step','fov_per_s')})
    else: print(l.rstrip()[:300])
"
done
