timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TKBO_PRECISION=fast timeout 250 python tools/stress_k1.py 60 9 2>&1 | tail -3
for w in c3s c4s c4d; do for pr in fast; do
TKBO_PRECISION=$pr timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-mpes --no-cpu 2>/dev/null | python tools/_pj.py $w $pr
done; done
TKBO_PRECISION=fast timeout 200 python bench.py --workload c2 --steps 50 --warmup 3 --no-mpes --no-cpu 2>/dev/null | python tools/_pj.py c2 fast
