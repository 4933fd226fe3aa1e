# GPU box: per-env logic duration statistics (reset vs normal steps) for a few games
for gm in ${GAMES:-"coinrun easy" "jumper hard" "caveflyer hard" "leaper hard" "bossfight hard"}; do
  set -- $gm
  echo "== $1 $2"; python tools/gpu_timing_hist.py $1 $2 32768 2>&1 | tail -4
done
