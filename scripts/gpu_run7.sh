for t in "dog_variant=1" "dog_variant=0,dog_ctas=1"; do
echo "== $t"
AMT_TRACE=1 AMT_TUNE="$t" python bench.py --fovs 32 --steps 1 --warmup 1 --no-e2e --no-cpu 2>&1 | grep amt-trace | tail -44
done
