mkdir -p gpurun_out
for sp in 1 0; do for g in "leaper hard 32768" "chaser hard 32768" "coinrun easy 65536" "jumper hard 32768" "starpilot hard 32768"; do set -- $g
PGB200_SPLIT_RESET=$sp python bench.py --game $1 --mode $2 --envs-per-gpu $3 --steps 30 --warmup 5 --desync-steps 400 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('split=$sp %-10s steady %7.2f M/s cold %7.2f step %6.3f ms logic %6.3f render %6.3f launches %d' % (j['config']['game'][:10], j['value']/1e6, j['value_cold']/1e6, j['ms_per_step'], r['logic_kernel_ms_avg'], r['kernel_ms_avg'], j['gpu_launches']))"
done; done
