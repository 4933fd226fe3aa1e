mkdir -p gpurun_out
for ch in 1 2 4 8 16; do
python bench.py --steps 40 --warmup 5 --desync-steps 500 --chunks $ch --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('chunks $ch coinrun steady %7.2f M/s step %6.3f' % (j['value']/1e6, j['ms_per_step']))"
python bench.py --game maze --mode hard --envs-per-gpu 32768 --steps 40 --warmup 5 --desync-steps 500 --chunks $ch --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('chunks $ch maze    steady %7.2f M/s step %6.3f' % (j['value']/1e6, j['ms_per_step']))"
done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
