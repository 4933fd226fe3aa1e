# GPU box: parity subset, then a sweep of the co-residency knobs (same library, env vars) on a few games
mkdir -p gpurun_out
: > gpurun_out/knobs.jsonl
if [ "${RUN_TESTS:-1}" = "1" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "libenv_host_buffers or sixteen" 2>&1 | tail -3; fi
KNOBS=${KNOBS:-0:24 7:24 6:24 5:24 0:12 6:12 6:16 5:16}
for k in $KNOBS; do
  IFS=: read R L P <<< "$k"
  export PGB200_RENDER_CTAS_PER_SM=$R PGB200_LOGIC_BLOCKS_PER_SM=$L PGB200_PRIORITY_SPLIT=${P:-0}
  for g in ${GAMES:-coinrun65k}; do
    m=hard; e=${GAME_ENVS:-32768}
    if [ "$g" = "coinrun65k" ]; then g=coinrun; m=easy; e=65536; fi
    timeout 300 python bench.py --game $g --mode $m --envs-per-gpu $e --steps 40 --warmup 8 --no-e2e --no-cpu-baseline 2>> gpurun_out/knobs.err | sed "s/^{/{\"variant\": \"R$PGB200_RENDER_CTAS_PER_SM-L$PGB200_LOGIC_BLOCKS_PER_SM-P$PGB200_PRIORITY_SPLIT\", /" >> gpurun_out/knobs.jsonl
  done
done
tail -3 gpurun_out/knobs.err
python - <<'PY'
import json
for l in open('gpurun_out/knobs.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    r=d['roofline']
    print('%-8s %-10s %5s %6.2f M/s step %6.3f ms | serial: logic %6.3f render %6.3f | err %s' % (d['variant'], d['config']['game'][:10], d['config']['distribution_mode'], d['value']/1e6, d['ms_per_step'], r['logic_kernel_ms_avg'], r['kernel_ms_avg'], d.get('env_error_bits')))
PY
