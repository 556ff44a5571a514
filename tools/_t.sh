timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rff_fused_kernel -s 3 -c 1 -o gpurun_out/prof_k1_c2_v8 -f python bench.py --steps 2 --warmup 3 --no-mpes --no-cpu > gpurun_out/ncu_c2.log 2>&1; tail -2 gpurun_out/ncu_c2.log | cut -c1-200
