"""Print the key figures of bench.py JSON lines: python tools/show_bench.py gpurun_out/r2p_*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        km = r.get('kernel_ms') or {}
        print('%-34s N=%d K=%-5d %8.3f us  %.3e %s  step %.3f  kern %s  e2e %.3f ms %.3e  cpu %s  traffic %s' % (
            f.split('/')[-1], d['n_gpus'], d['steps'], d['ms_per_step'] * 1e3, d['value'], d['unit'].split('/')[0][:5],
            (r.get('step') or {}).get('frac', float('nan')), {k: round(v * 1e3, 2) for k, v in km.items()},
            d['e2e'].get('ms_per_step', float('nan')), d['e2e']['value'],
            ('%.3e' % d['cpu_baseline']['value']) if d.get('cpu_baseline') else None, r.get('traffic')))
    except Exception as e:  # noqa
        print('%-34s ERR %s' % (f.split('/')[-1], e))
