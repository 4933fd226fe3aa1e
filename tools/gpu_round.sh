# Run on the GPU box via gpurun: tests, bench, ncu launch list + one full capture of the step kernel.
mkdir -p gpurun_out
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ "${SKIP_TESTS:-0}" != "1" ]; then timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
python bench.py --steps ${BENCH_STEPS:-100} --warmup 10 --e2e-steps 5 --cpu-budget 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"logic_kernel|render_kernel" -s 8 -c 2 -f -o gpurun_out/prof_step \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
fi
