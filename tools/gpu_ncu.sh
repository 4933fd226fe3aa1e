mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"logic_kernel|render_kernel" -s 40 -c 2 -f -o gpurun_out/prof_step \
    python bench.py --steps 25 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
