"""Per-source-line instruction counts / stall samples of one kernel from an ncu report captured with --import-source on.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, hdr, out, tot_i, tot_s = None, None, [], 0, 0
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) > 6 and r[0] == 'Line No':
        hdr = r
        ci, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= max(ci, si) or r[2] != '-':
        continue  # SASS rows carry an address in column 2; source rows have '-'
    try:
        n, s = int(r[ci]), int(r[si])
    except ValueError:
        continue
    tot_i += n
    tot_s += s
    out.append((n, s, cur_file, r[0], r[1].strip()[:100]))
print('total warp instructions %d, samples %d' % (tot_i, tot_s))
for n, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print('%9d %5.1f%%  samp %5.1f%%  %s:%s | %s' % (n, 100.0 * n / max(tot_i, 1), 100.0 * s / max(tot_s, 1), f, ln, src))
