# render phases, logic A/B, reset costs
mkdir -p gpurun_out
PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_phase.so python tools/gpu_render_phases.py coinrun easy 65536 600 2>&1 | tail -10 | tee gpurun_out/phases_coinrun.txt
PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_phase.so python tools/gpu_render_phases.py maze hard 32768 400 2>&1 | tail -10 | tee gpurun_out/phases_maze.txt
bash tools/gpu_ab_logic.sh 2>&1 | grep -v "^==" 
python tools/gpu_reset_cost.py 32768 500 2>&1 | tee gpurun_out/reset_cost.jsonl | cut -c1-400
