"""Print a compact set of metrics from .ncu-rep files:  python tools/ncu_keys.py rep [rep ...]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_uniform.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warps_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warps_issue_stalled_membar_per_warp_active.pct', 'smsp__warps_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warps_issue_stalled_sleeping_per_warp_active.pct', 'smsp__warps_issue_stalled_wait_per_warp_active.pct',
        'smsp__warps_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warps_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warps_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warps_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warps_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warps_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warps_issue_stalled_tex_throttle_per_warp_active.pct', 'smsp__warps_issue_stalled_selected_per_warp_active.pct',
        'smsp__warps_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warps_issue_stalled_drain_per_warp_active.pct',
        'smsp__warps_issue_stalled_imc_miss_per_warp_active.pct', 'smsp__warps_issue_stalled_misc_per_warp_active.pct']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', rep, r[hdr.index('Kernel Name')][:60], 'grid', r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
        for k in WANT:
            if k in hdr:
                print(f"  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}")
