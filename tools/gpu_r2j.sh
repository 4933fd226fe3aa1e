mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "(libenv_host_buffers and (coinrun or maze-hard or heist-hard or dodgeball-hard or fruitbot or bossfight or jumper-easy or starpilot-hard or miner or chaser)) or non_default or unsnapped" 2>&1 | tail -4
PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_phase.so python tools/gpu_render_phases.py coinrun easy 65536 600 2>&1 | tail -10 | tee gpurun_out/phases_coinrun.txt
python bench.py --steps 60 --warmup 5 --desync-steps 1000 --e2e-steps 5 --cpu-budget 5 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -3 gpurun_out/bench_j.err
bash tools/gpu_allgames_r2.sh 2>&1 | tail -20
