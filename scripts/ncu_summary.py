#!/usr/bin/env python
"""Summarises an .ncu-rep (one `ncu --set full` capture) into a small JSON + the top stall lines, for profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_fused_r1.ncu-rep profiles/r1_fused.json [kernel-regex]

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'dram__bytes_write.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_tensor.sum', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'smsp__cycles_active.avg', 'gpc__cycles_elapsed.max',
]


def ncu_csv(report, page, extra=()):
    out = subprocess.run(['ncu', '-i', report, '--page', page, '--csv'] + list(extra), capture_output=True, text=True)
    return list(csv.reader(io.StringIO(out.stdout)))


def main():
    report, dest = sys.argv[1], sys.argv[2]
    rows = ncu_csv(report, 'raw')
    hdr, units = rows[0], rows[1]
    summary = {'report': report, 'kernels': []}
    for r in rows[2:]:
        k = {'name': r[hdr.index('Kernel Name')]}
        for i, h in enumerate(hdr):
            base = h.split('.TriageCompute.')[-1]
            if base in METRICS:
                try:
                    k[base] = {'value': float(r[i].replace(',', '')), 'unit': units[i]}
                except ValueError:
                    pass
        summary['kernels'].append(k)
    # top stall sites
    src = ncu_csv(report, 'source')
    tops = []
    hdr_i = next((i for i, r in enumerate(src) if 'Source' in r and 'Address' in r), None)
    if hdr_i is not None:
        h = src[hdr_i]
        si, so = h.index('Warp Stall Sampling (All Samples)'), h.index('Source')
        stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
        data = [r for r in src[hdr_i + 1:] if len(r) > si and r[si].isdigit()]
        seen = set()
        total = sum(int(r[si]) for r in data) or 1
        for r in sorted(data, key=lambda r: -int(r[si])):
            key = (r[0], r[so])
            if key in seen:
                continue
            seen.add(key)
            st = sorted(((h[i], int(r[i])) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0),
                        key=lambda x: -x[1])[:2]
            tops.append({'samples_pct': round(100.0 * int(r[si]) / total, 2), 'sass': r[so].strip(), 'stalls': st})
            if len(tops) >= 12:
                break
    summary['top_stall_sites'] = tops
    json.dump(summary, open(dest, 'w'), indent=1)
    for k in summary['kernels']:
        print(k['name'][:90])
        for m, v in k.items():
            if m != 'name':
                print('   %-75s %s %s' % (m, v['value'], v['unit']))
    for t in tops:
        print('  %5.2f%%  %-60s %s' % (t['samples_pct'], t['sass'][:60], t['stalls']))


if __name__ == '__main__':
    main()
