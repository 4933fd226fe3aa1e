mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err
bash tools/gpu_allgames_r2.sh 2>&1 | tail -20
