mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --desync-steps 300 --e2e-steps 5 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -3 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json | cut -c1-3000
