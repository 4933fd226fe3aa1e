# A/B the logic kernel's warps-per-CTA / register cap on the GPU box (rebuilds there).
mkdir -p gpurun_out
for cfg in "1 1" "2 16" "2 20" "2 24" "4 12" "2 32"; do
  set -- $cfg
  PG_NVCC_EXTRA="-DPG_LOGIC_WARPS=$1 -DPG_LOGIC_MIN_BLOCKS=$2" python -c "from procgen_b200 import build as B; B.build_library(force=True)" 2>&1 | grep -i error
  echo "cfg warps=$1 minblocks=$2"
  python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
