for prio in -1 0; do
for t in "dog_variant=0,dog_ctas=2,stream_ctas=16" "dog_variant=0,dog_ctas=2,stream_ctas=16,stream_pad_kb=100" "dog_variant=1,stream_ctas=16" "dog_variant=0,dog_ctas=2,stream_ctas=2"; do
  echo "== B priority $prio, $t"; PROBE_B_PRIORITY=$prio AMT_TUNE="$t" python scripts/overlap_probe.py 2>&1 | tail -1
done; done
