mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "(libenv_host_buffers and (coinrun or maze-hard or chaser or miner)) or non_default" 2>&1 | tail -3
python bench.py --steps 60 --warmup 5 --desync-steps 1000 --e2e-steps 5 --cpu-budget 5 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -3 gpurun_out/bench_k.err
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|setup_kernel|logic_kernel" -s 2145 -c 3 -f -o gpurun_out/prof_step_k \
    python bench.py --steps 5 --warmup 3 --desync-steps 700 --chunks 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
