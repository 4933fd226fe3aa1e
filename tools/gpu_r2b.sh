# bench + ncu list + full capture of render and logic kernels at steady state (chunks=1)
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "libenv_host_buffers and (coinrun or maze-hard or heist-hard or chaser or fruitbot or bossfight)" 2>&1 | tail -4
python bench.py --steps 60 --warmup 5 --desync-steps ${DESYNC:-1000} --e2e-steps 5 --cpu-budget 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; cat gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|logic_kernel" -s 1430 -c 2 -f -o gpurun_out/prof_step_b \
    python bench.py --steps 5 --warmup 3 --desync-steps 700 --chunks 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
