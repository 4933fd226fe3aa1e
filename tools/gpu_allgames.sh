# GPU box: parity tests, headline bench, one bench line per game, ncu launch list + full capture.
mkdir -p gpurun_out
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps ${BENCH_STEPS:-100} --warmup 10 --e2e-steps 5 --cpu-budget 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
: > gpurun_out/bench_games.jsonl
for g in bigfish bossfight caveflyer chaser climber coinrun dodgeball fruitbot heist jumper leaper maze miner ninja plunder starpilot; do
  timeout 300 python bench.py --game $g --mode hard --envs-per-gpu ${GAME_ENVS:-32768} --steps 60 --warmup 10 --no-e2e --no-cpu-baseline >> gpurun_out/bench_games.jsonl 2>> gpurun_out/bench_games.err
done
cat gpurun_out/bench_games.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['game'], d['config']['distribution_mode'], round(d['value'] / 1e6, 2), 'M steps/s', d['ms_per_step'], 'ms', 'err', d.get('env_error_bits'))
"
if [ "${SKIP_NCU:-0}" != "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"logic_kernel|render_kernel" -s 40 -c 4 -f -o gpurun_out/prof_step \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
fi
