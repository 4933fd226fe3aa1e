# GPU box: quick record check — smoke, every GPU test, headline bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 10 --e2e-steps 8 --cpu-budget 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value', round(d['value']/1e6,2), 'M/s  step', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']/1e6,2), 'render', round(d['roofline']['kernel_ms_avg'],3), 'logic', round(d['roofline']['logic_kernel_ms_avg'],3), 'frac', round(d['roofline']['frac'],4), 'err', d['env_error_bits'])"
for g in bossfight fruitbot bigfish; do python bench.py --game $g --mode hard --envs-per-gpu 32768 --steps 40 --warmup 8 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['config']['game'], round(d['value']/1e6,2),'M/s render', round(r['kernel_ms_avg'],3), 'logic', round(r['logic_kernel_ms_avg'],3), 'err', d['env_error_bits'])"; done
