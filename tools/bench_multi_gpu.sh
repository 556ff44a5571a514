#!/bin/bash
# bench.py on N GPUs of this node for the given workloads, one digest line each (results in gpurun_out/):
#   bash tools/bench_multi_gpu.sh 8 c2 c3s c4d
mkdir -p gpurun_out
N=$1; shift
for w in "$@"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --no-mpes --no-cpu > gpurun_out/mg${N}_$w.json 2> gpurun_out/mg${N}_$w.err
python tools/bench_line.py $w x$N < gpurun_out/mg${N}_$w.json | cut -c1-120
python -c "
import json
j=json.loads(open('gpurun_out/mg${N}_$w.json').read().strip().splitlines()[-1])
print('   e2e ms', round(j['e2e']['ms_per_step'],3), 'e2e value %.3e' % j['e2e']['value'], '|', j['config']['parallelism'][:90])
"
done
