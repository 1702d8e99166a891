for s in none SSDN_CONV_PAIR=0 SSDN_PDL=0 SSDN_WGRAD_STREAM=0 SSDN_WGRAD_NINE=0 SSDN_EPI_SPLIT=0; do
  echo "== $s"
  if [ "$s" == "none" ]; then timeout 100 python tests/dev_small_batch.py 4 2>/dev/null | tail -1; else env $s timeout 100 python tests/dev_small_batch.py 4 2>/dev/null | tail -1; fi
done
