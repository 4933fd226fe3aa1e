export PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_phase.so
PG_SLOWEST=1 python tools/gpu_timing_hist.py coinrun easy 65536 2>&1 | tail -18
