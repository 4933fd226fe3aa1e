mkdir -p gpurun_out
for v in "" _sb4 _sb6 _sb8; do
  lib=$PWD/procgen_b200/libprocgen_b200$v.so
  for g in "coinrun easy 65536" "maze hard 32768" "bossfight hard 32768" "fruitbot hard 32768"; do
    set -- $g
    PROCGEN_B200_LIB=$lib python bench.py --game $1 --mode $2 --envs-per-gpu $3 --steps 30 --warmup 5 --desync-steps 300 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); r=j['roofline']
print('variant %-5s %-10s steady %7.2f M/s step %6.3f logic %6.3f setup %6.3f render %6.3f' % ('$v', '$1', j['value']/1e6, j['ms_per_step'], r['logic_kernel_ms_avg'], r['setup_kernel_ms_avg'], r['kernel_ms_avg']))" | tee -a gpurun_out/ab_setup.txt
  done
done
