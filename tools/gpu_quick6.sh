timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "libenv_host_buffers and (coinrun or bigfish or maze-hard or bossfight or miner or climber)" 2>&1 | tail -2
for g in "coinrun easy 65536" "bigfish hard 65536" "maze hard 32768" "bossfight hard 32768" "climber hard 32768"; do set -- $g
python bench.py --game $1 --mode $2 --envs-per-gpu $3 --steps 40 --warmup 5 --desync-steps 500 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('%-10s steady %7.2f M/s cold %7.2f step %6.3f logic %6.3f setup %6.3f render %6.3f err %d' % (j['config']['game'][:10], j['value']/1e6, j['value_cold']/1e6, j['ms_per_step'], r['logic_kernel_ms_avg'], r['setup_kernel_ms_avg'], r['kernel_ms_avg'], j['env_error_bits']))"
done
