mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --desync-steps 300 --e2e-steps 5 --cpu-budget 5 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; tail -3 gpurun_out/bench_4gpu.err; grep '^{' gpurun_out/bench_4gpu.json | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('N=4 value %.2f M/s step %.3f ms e2e %.2f M/s (%.1f GB/s/rank) clocks %s' % (j['value']/1e6, j['ms_per_step'], j['e2e']['value']/1e6, j['e2e']['d2h_gbs_per_rank'], j['clocks']))
print(json.dumps(j.get('config5'), indent=1))"
