for cfg in "ru_wc=2 ru_cpt=3 ru_waves=8" "ru_wc=2 ru_cpt=3 ru_waves=1" "ru_wc=4 ru_cpt=3 ru_waves=1" "ru_wc=4 ru_cpt=3 ru_waves=8" "ru_wc=4 ru_cpt=1 ru_waves=8" "ru_wc=2 ru_cpt=3 ru_waves=2"; do
  args=""; for kv in $cfg; do args="$args --opt $kv"; done
  echo "== $cfg"
  CEIT_REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rank_update python tools/stage_breakdown.py 400 9 $args 2>&1 | grep -E "gpu__time_duration|excitation" | tail -3
done
