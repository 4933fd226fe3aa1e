"""Key per-kernel metrics from an ncu report. usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warps_issue_stalled', 'smsp__pcsamp_warps_issue_stalled', 'sm__cycles_elapsed.max', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'smsp__cycles_active.avg', 'sm__inst_executed_pipe_fp64', 'smsp__inst_executed_pipe_fp64']
for r in rows[2:]:
    print('===', r[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in keys) and r[i] not in ('0', '', 'n/a'):
            if 'pcsamp' in h and float(r[i].replace(',', '')) < 1: continue
            print(f'  {h} [{units[i]}] {r[i]}')
