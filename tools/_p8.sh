timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 250 python tools/stress_k1.py 80 10 2>&1 | tail -3
for w in c2 c3s c4s; do
timeout 200 python bench.py --workload $w --steps 20 --warmup 3 --no-mpes --no-cpu 2>/dev/null | python tools/_pj.py $w auto
done
TKBO_PRECISION=fp32 timeout 200 python bench.py --workload c3s --steps 10 --warmup 3 --no-mpes --no-cpu 2>/dev/null | python tools/_pj.py c3s fp32
timeout 120 python tools/trace_k1.py 151552 4096 6 256 100 2 2>&1 | grep -A8 "^items"
