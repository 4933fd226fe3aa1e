# GPU box: the round's record run — smoke, every GPU test, headline bench (with e2e and CPU reference), reference arm,
# one bench line per game, ncu launch list and full captures of both step kernels.
mkdir -p gpurun_out
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 100 --warmup 10 --e2e-steps 8 --cpu-budget 12 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 8 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_reference.json
bash tools/gpu_games.sh
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"logic_kernel|render_kernel" -s 40 -c 4 -f -o gpurun_out/prof_step \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
fi
