mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 3 --e2e-steps 3 --cpu-budget 5 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; tail -c 3000 gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
