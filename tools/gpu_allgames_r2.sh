# steady-state bench of every game (32768 envs, hard) + the BASELINE configs; one JSON line each
mkdir -p gpurun_out
out=gpurun_out/bench_all_games.jsonl; : > $out
run() { python bench.py --game $1 --mode $2 --envs-per-gpu $3 --steps ${4:-30} --warmup 5 --desync-steps ${5:-600} --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 >> $out; tail -1 $out | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('%-10s %-5s %6d  steady %7.2f M/s  cold %7.2f M/s  step %6.3f ms  logic %6.3f  render %6.3f  frac %.4f  ends/step %.5f err %d' % (j['config']['game'][:10], j['config']['distribution_mode'], j['config']['envs_per_gpu'], j['value']/1e6, j['value_cold']/1e6, j['ms_per_step'], r['logic_kernel_ms_avg'], r['kernel_ms_avg'], r['frac'], j['steady_state']['episode_end_fraction_per_step'], j['env_error_bits']))"; }
run coinrun easy 65536 60 1000
run bigfish hard 65536 40 600
run maze hard 32768
run heist hard 32768
for g in bigfish bossfight caveflyer chaser climber coinrun dodgeball fruitbot jumper leaper miner ninja plunder starpilot; do run $g hard 32768; done
run all16 hard 32768 30 300
