"""Record the DRAM traffic of a kernel from an ncu report into profiles/ncu_traffic.json, tied to the CUDA sources it was
built from (bench.py refuses the entry when those sources have changed since).

    python tools/update_traffic.py <bench kernel key> <report.ncu-rep> <source.cu> [more sources...]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    key, rep, sources = sys.argv[1], sys.argv[2], sys.argv[3:]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(name):
        v, u = float(r[idx[name]].replace(',', '')), units[idx[name]]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]

    path = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        data = json.load(open(path))
    except (OSError, ValueError):
        data = {}
    data[key] = {'kernel': r[idx['Kernel Name']], 'dram_read': val('dram__bytes_read.sum'),
                 'dram_write': val('dram__bytes_write.sum'), 'capture': os.environ.get('SUMMARY', os.path.basename(rep)), 'sources': sources,
                 'source_sha16': bench.kernel_source_sha(sources),
                 'note': 'one launch, ncu --set full --clock-control none; outputs still dirty in the 126 MB L2 at kernel end '
                         'are not counted as DRAM writes'}
    json.dump(data, open(path, 'w'), indent=1, sort_keys=True)
    print(key, data[key])


if __name__ == '__main__':
    main()
