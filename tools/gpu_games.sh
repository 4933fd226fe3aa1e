# GPU box: one bench line per game with the per-kernel split (no tests)
mkdir -p gpurun_out
: > gpurun_out/bench_games.jsonl
for g in ${GAMES:-bigfish bossfight caveflyer chaser climber coinrun dodgeball fruitbot heist jumper leaper maze miner ninja plunder starpilot all16}; do
  timeout 300 python bench.py --game $g --mode hard --envs-per-gpu ${GAME_ENVS:-32768} --steps 40 --warmup 8 --no-e2e --no-cpu-baseline >> gpurun_out/bench_games.jsonl 2>> gpurun_out/bench_games.err
done
tail -3 gpurun_out/bench_games.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_games.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print('%-10s %6.2f M/s step %6.3f ms | serial: logic %6.3f ms render %6.3f ms (x%d launches) | render roofline %.4f err %s' % (d['config']['game'][:10], d['value']/1e6, d['ms_per_step'], r['logic_kernel_ms_avg'], r['kernel_ms_avg'], r['launches_timed'], r['frac'], d.get('env_error_bits')))
PY
