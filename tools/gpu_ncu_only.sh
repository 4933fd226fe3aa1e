mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"logic_kernel|render_kernel" -s 40 -c 4 -f -o gpurun_out/prof_step \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -4
