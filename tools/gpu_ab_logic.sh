# A/B of logic-kernel register budgets (build variants next to the product library)
mkdir -p gpurun_out
for v in "" _lb16 _lb12w4 _lb32; do
  lib=$PWD/procgen_b200/libprocgen_b200$v.so
  for g in "coinrun easy 65536" "maze hard 32768"; do
    set -- $g
    echo "== variant '$v' $1"
    PROCGEN_B200_LIB=$lib python bench.py --game $1 --mode $2 --envs-per-gpu $3 --steps 40 --warmup 5 --desync-steps 400 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print(json.dumps({'variant':'$v','game':'$1','value':j['value'],'value_cold':j['value_cold'],'ms':j['ms_per_step'],'logic_ms':r['logic_kernel_ms_avg'],'render_ms':r['kernel_ms_avg'],'err':j['env_error_bits'],'launches':j['gpu_launches']}))" | tee -a gpurun_out/ab_logic.jsonl
  done
done
