mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|setup_kernel|logic_kernel" -s 2145 -c 3 -f -o gpurun_out/prof_step_final \
    python bench.py --steps 5 --warmup 3 --desync-steps 700 --chunks 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
bash tools/gpu_allgames_r2.sh 2>&1 | tail -20
