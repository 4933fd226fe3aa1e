# last verification of the round: the whole GPU suite, smoke(), and the default bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py --steps 100 --warmup 10 --cpu-budget 8 > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err; tail -2 gpurun_out/bench_verify.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_verify.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
PY
