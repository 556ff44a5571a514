N=$1
for w in c2 c3s c4d; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps 20 --warmup 3 --no-mpes --no-cpu > gpurun_out/mg${N}_$w.json 2> gpurun_out/mg${N}_$w.err
python tools/_pj.py $w x$N < gpurun_out/mg${N}_$w.json
done
