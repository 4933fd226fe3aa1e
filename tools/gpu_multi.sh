# GPU box with N GPUs (gpurun --gpus N): scaling of the headline config and the 16-game list, with/without gather
N=${NGPU:-2}
mkdir -p gpurun_out
set -x
: > gpurun_out/scale.jsonl
python bench.py --gpus 1 --steps 60 --warmup 10 --no-e2e --no-cpu-baseline >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 10 --e2e-steps 5 --no-cpu-baseline >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --game all16 --mode hard --envs-per-gpu 32768 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --game all16 --mode hard --envs-per-gpu 32768 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --gather >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
python bench.py --gpus 1 --game all16 --mode hard --envs-per-gpu 32768 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline >> gpurun_out/scale.jsonl 2>> gpurun_out/scale.err
tail -20 gpurun_out/scale.err
python - <<'PY'
import json
for l in open('gpurun_out/scale.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d['n_gpus'], d['config']['game'][:30], d['config']['parallelism'], round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step'],3),'ms', 'e2e', (d.get('e2e') or {}).get('value'), 'roofline', round(d['roofline']['frac'],4), d['roofline']['kernel_ms_avg'])
PY
