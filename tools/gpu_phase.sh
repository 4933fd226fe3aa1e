# GPU box: parity subset on the default build, phase timing on the profiling variant, one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "libenv_host_buffers or sixteen or state" 2>&1 | tail -4
python bench.py --steps 40 --warmup 8 --e2e-steps 8 --no-cpu-baseline 2>gpurun_out/b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('coinrun easy 65536:', round(d['value']/1e6, 2), 'M/s  e2e', round(d['e2e']['value']/1e6, 2), 'M/s  render ms', d['roofline']['kernel_ms_avg'], 'frac', d['roofline']['frac'])"
export PROCGEN_B200_LIB=$PWD/procgen_b200/libprocgen_b200_phase.so
for gm in "jumper hard" "leaper hard" "caveflyer hard"; do set -- $gm; echo "== $1 $2"; PG_PHASES=1 python tools/gpu_timing_hist.py $1 $2 32768 2>&1 | tail -5; done
unset PROCGEN_B200_LIB
ncu --set full --clock-control none --import-source on -k regex:"render_kernel" -s 20 -c 2 -f -o gpurun_out/prof_dodgeball \
    python bench.py --game dodgeball --mode hard --envs-per-gpu 32768 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_dodge.log 2>&1
ls -la gpurun_out | tail -4
