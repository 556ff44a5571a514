timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c2 c3s; do timeout 200 python bench.py --workload $w --no-mpes --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$w', 'kernel ms', round(j['ms_per_step'],3), 'e2e ms', round(j['e2e']['ms_per_step'],3), 'e2e value %.3e' % j['e2e']['value'])
"; done
