#!/bin/bash
# BASELINE.json configs[4]: the C5 H(e,e'K+)Lambda sweep, 1e11 generated tries in total on N GPUs of one box:
#   tools/c5_sweep.sh N   -> gpurun_out/r2_c5_1e11_n<N>.json  (device-timed steps only: bench.py --skip-e2e)
# 745 steps of 2^27 tries = 1.0e11; with N ranks each rank runs ceil(745 / N) steps of its own try ranges.
mkdir -p gpurun_out
N=$1
STEPS=$(( (745 + N - 1) / N ))
OUT=gpurun_out/r2_c5_1e11_n$N
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --config c5 --steps $STEPS --warmup 3 --tries 134217728 --skip-e2e --no-cpu-baseline > $OUT.json 2> $OUT.err < /dev/null
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config c5 --steps $STEPS --warmup 3 --tries 134217728 --skip-e2e --no-cpu-baseline > $OUT.json 2> $OUT.err < /dev/null
fi
tail -c 400 $OUT.err
python - <<PY
import json
for line in open("$OUT.json"):
    try: d = json.loads(line)
    except Exception: continue
    print("C5 N=$N: %.4g generated ev/s, %d steps x %.1f ms = %.1f s for %.3g tries" % (d["value"], d["steps"], d["ms_per_step"], d["steps"] * d["ms_per_step"] / 1e3, d["steps"] * d["config"]["tries_per_step_per_gpu"] * d["n_gpus"]))
PY
