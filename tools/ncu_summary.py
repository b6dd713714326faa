"""Turn ncu artefacts from gpurun_out/ into the small tracked summaries under profiles/.
   python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches_summary.csv
   python tools/ncu_summary.py full gpurun_out/prof_conv_tc_r01.ncu-rep profiles/r01_conv_tc_full.csv"""
import collections
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    tot = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u in ('nsecond', 'ns') else v * 1e3 if u in ('msecond', 'ms') else v
        k = row['Kernel Name'].split('(')[0]
        tot.setdefault(k, [0, 0.0])
        tot[k][0] += 1
        tot[k][1] += v
    T = sum(v[1] for v in tot.values())
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_us', 'share_pct', 'avg_us'])
        for k, (n, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
            w.writerow([k, n, round(t, 1), round(100 * t / T, 2), round(t / n, 2)])
        w.writerow(['TOTAL', sum(v[0] for v in tot.values()), round(T, 1), 100.0, ''])


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow([f"{k} [{units[i]}]" if units[i] else k for k, i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:80] for _, i in idx])


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
