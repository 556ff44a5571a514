set -x
timeout 600 python bench.py > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err
for w in c3s c4s c4d; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-mpes > gpurun_out/final_$w.json 2> gpurun_out/final_$w.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_v8.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
