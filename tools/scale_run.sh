#!/bin/bash
# One line of the weak-scaling series: tools/scale_run.sh N [config] -> gpurun_out/scale_<config>_n<N>.json
mkdir -p gpurun_out
N=$1; C=${2:-c1}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $C --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${C}_n$N.json 2> gpurun_out/scale_${C}_n$N.err
python tools/bench_line.py ${C}_n$N < gpurun_out/scale_${C}_n$N.json
