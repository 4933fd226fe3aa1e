timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "libenv_host_buffers and (coinrun or bossfight or fruitbot or heist-hard or dodgeball-hard or starpilot-hard)" 2>&1 | tail -2
for v in "" _noinl _lb16 _lb20; do
for g in "coinrun easy 65536" "bossfight hard 32768" "fruitbot hard 32768"; do set -- $g
PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200$v.so python bench.py --game $1 --mode $2 --envs-per-gpu $3 --steps 40 --warmup 5 --desync-steps 400 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('variant %-6s %-10s steady %7.2f M/s step %6.3f logic %6.3f setup %6.3f render %6.3f err %d' % ('$v', j['config']['game'][:10], j['value']/1e6, j['ms_per_step'], r['logic_kernel_ms_avg'], r['setup_kernel_ms_avg'], r['kernel_ms_avg'], j['env_error_bits']))"
done; done
