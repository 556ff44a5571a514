TKBO_ACC1=1 timeout 200 python bench.py --no-mpes --no-cpu 2>/dev/null | python tools/_pj.py c2 acc1-ring7
timeout 200 python bench.py --no-mpes --no-cpu 2>/dev/null | python tools/_pj.py c2 base
TKBO_ACC1=1 timeout 100 python tools/stress_k1.py 12 13 2>&1 | tail -3
