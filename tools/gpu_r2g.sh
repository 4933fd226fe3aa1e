mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "libenv_host_buffers and (coinrun or maze-hard or heist-hard or dodgeball-hard or fruitbot or bossfight or jumper-easy or starpilot-hard or miner)" 2>&1 | tail -4
PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_phase.so python tools/gpu_render_phases.py coinrun easy 65536 600 2>&1 | tail -10 | tee gpurun_out/phases_coinrun.txt
python bench.py --steps 60 --warmup 5 --desync-steps 1000 --e2e-steps 5 --cpu-budget 5 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -3 gpurun_out/bench_g.err
ncu --set full --clock-control none --import-source on -k regex:"render_kernel|setup_kernel|logic_kernel" -s 2145 -c 3 -f -o gpurun_out/prof_step_g \
    python bench.py --steps 5 --warmup 3 --desync-steps 700 --chunks 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
