"""Summarise an ncu report (run in the build container: ncu reads .ncu-rep files without a GPU).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md
"""
import csv
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('launch__shared_mem_per_block_static', 'static smem/block'),
    ('launch__occupancy_limit_registers', 'occupancy limit (regs) blocks/SM'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem) blocks/SM'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('sm__cycles_elapsed.max', 'SM cycles elapsed'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('# ncu summary of `%s`\n' % path)
    print('Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; per-launch '
          'values (cold caches, serialised launches: compare shares, not absolutes).\n')
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        print('## `%s`\n' % name[:110])
        print('| metric | value |')
        print('|---|---|')
        for key, label in WANT:
            if key in idx and r[idx[key]] not in ('', 'n/a'):
                print('| %s | %s %s |' % (label, r[idx[key]], units[idx[key]]))
        stalls = []
        for h in hdr:
            if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
                v = r[idx[h]].replace(',', '')
                if v not in ('', 'n/a'):
                    stalls.append((h.replace('smsp__pcsamp_warps_issue_stalled_', ''), float(v)))
        tot = sum(v for _, v in stalls) or 1.0
        top = sorted(stalls, key=lambda x: -x[1])[:6]
        print('| top stall reasons (PC samples) | %s |' % ', '.join('%s %.0f%%' % (n, 100 * v / tot) for n, v in top))
        rd = float(r[idx['dram__bytes_read.sum']].replace(',', '')) if 'dram__bytes_read.sum' in idx else 0
        print()


if __name__ == '__main__':
    main(sys.argv[1])
