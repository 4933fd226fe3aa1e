# GPU box: A/B of library variants on a few games (no rebuild on the box). VARIANTS="default mb8 ..."
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
if [ "${RUN_TESTS:-1}" = "1" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "libenv_host_buffers or sixteen" 2>&1 | tail -4; fi
for v in ${VARIANTS:-default}; do
  if [ "$v" = "default" ]; then unset PROCGEN_B200_LIB; else export PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_$v.so; fi
  for g in ${GAMES:-coinrun bossfight dodgeball bigfish}; do
    m=hard; e=${GAME_ENVS:-32768}
    if [ "$g" = "coinrun65k" ]; then g=coinrun; m=easy; e=65536; fi
    timeout 300 python bench.py --game $g --mode $m --envs-per-gpu $e --steps 40 --warmup 8 --no-e2e --no-cpu-baseline 2>> gpurun_out/ab.err | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/ab.jsonl
  done
done
tail -3 gpurun_out/ab.err
python - <<'PY'
import json
for l in open('gpurun_out/ab.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print('%-8s %-10s %5s %6.2f M/s step %6.3f ms | serial: logic %6.3f render %6.3f | err %s' % (d['variant'], d['config']['game'][:10], d['config']['distribution_mode'], d['value']/1e6, d['ms_per_step'], r['logic_kernel_ms_avg'], r['kernel_ms_avg'], d.get('env_error_bits')))
PY
