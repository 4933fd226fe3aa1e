# round-2 final evidence run: parity, headline bench, ncu of the three step kernels, all-games table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "libenv_host_buffers or sixteen or non_default or unsnapped or mid_array" 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|setup_kernel|logic_kernel" -s 2145 -c 3 -f -o gpurun_out/prof_step_final \
    python bench.py --steps 5 --warmup 3 --desync-steps 700 --chunks 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 96 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 5 --warmup 3 --desync-steps 400 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
bash tools/gpu_allgames_r2.sh 2>&1 | tail -20
