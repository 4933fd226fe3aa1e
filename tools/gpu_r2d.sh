mkdir -p gpurun_out
timeout 1400 python -m pytest tests/test_gpu_parity.py -q -x -k "libenv_host_buffers or state_blobs or golden or sixteen or remaining or non_default or unsnapped or consumer or wrappers" 2>&1 | tail -15
python tools/gpu_reset_cost.py 32768 500 dodgeball:hard fruitbot:hard heist:hard jumper:hard leaper:hard maze:hard miner:hard ninja:hard plunder:hard starpilot:hard 2>&1 | tee -a gpurun_out/reset_cost.jsonl | cut -c1-330
