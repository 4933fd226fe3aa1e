# round 2, first GPU call: smoke, quick parity subset, steady-state bench, ncu list + full capture
mkdir -p gpurun_out
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "libenv_host_buffers and (coinrun or maze or heist or chaser or jumper-easy or starpilot-hard or fruitbot or bossfight)" 2>&1 | tail -8
python bench.py --steps 60 --warmup 5 --e2e-steps 5 --cpu-budget 8 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 5700 -c 64 --csv --log-file gpurun_out/launches_a.csv \
    python bench.py --steps 5 --warmup 3 --desync-steps 700 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"render_kernel" -s 715 -c 1 -f -o gpurun_out/prof_render_a \
    python bench.py --steps 5 --warmup 3 --desync-steps 700 --chunks 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
