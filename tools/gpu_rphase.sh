# GPU box: render-kernel phase cycles (profiling variant) for a few games
export PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_phase.so
for gm in ${GAMES:-"coinrun easy" "dodgeball hard" "fruitbot hard" "bossfight hard"}; do set -- $gm; echo "== $1 $2"; PG_PHASES=1 PG_PHASES_ALL=1 python tools/gpu_timing_hist.py $1 $2 32768 2>&1 | tail -2; done
