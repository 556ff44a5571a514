#!/usr/bin/env python
"""One-line digest of a bench.py JSON line read from stdin:  python bench.py ... | python tools/bench_line.py [label ...]"""
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    t = j['config'].get('tiling', {})
    print(sys.argv[1:], round(j['ms_per_step'], 3), '%.3e' % j['value'], 'frac', round(j.get('roofline', {}).get('frac', 0), 3), 'stages', t.get('stages'), 'w', t.get('w_stages'), 'ring', t.get('tmem_ring'), 'c8', t.get('e4m3_corrections'), j.get('clocks'))
