mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1g_bench_n$N.json 2> gpurun_out/r1g_bench_n$N.err
python tools/bench_line.py n$N < gpurun_out/r1g_bench_n$N.json
