for t in "dog_variant=1,dog_ctas=1,stream_ctas=4" "dog_variant=1,dog_ctas=1,stream_ctas=6" "dog_variant=1,dog_ctas=1,stream_ctas=2" "dog_variant=1,dog_ctas=0,stream_ctas=1" "dog_variant=2,dog_ctas=1,stream_ctas=4" "dog_variant=0,dog_ctas=2,stream_ctas=2"; do
  echo "== $t"; AMT_TUNE="$t" python scripts/overlap_probe.py 2>&1 | tail -3
done
