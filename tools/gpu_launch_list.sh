# per-launch durations of the three step kernels (serialised, cold-cache: shares, not absolutes)
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 96 --csv --log-file gpurun_out/launches_verify.csv \
    python bench.py --steps 5 --warmup 3 --desync-steps 400 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches_verify.csv
