mkdir -p gpurun_out
for cfg in "1 1" "2 2" "4 4" "8 8"; do
  set -- $cfg
  PG_NVCC_EXTRA="-DPG_STEP_CHUNKS=$1 -DPG_AUX_STREAMS=$2" python -c "from procgen_b200 import build as B; B.build_library(force=True)" 2>&1 | grep -i " error"
  echo "cfg chunks=$1 streams=$2"
  python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['env_error_bits'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
grep -E "logic_kernel|render_kernel" gpurun_out/launches.csv | tail -8 | awk -F'","' '{print $5, $NF}' | cut -c1-120
