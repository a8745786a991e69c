#!/bin/bash
# One multi-GPU measurement call: tools/scale_run.sh N  ->  gpurun_out/r2_scale_c1_n<N>.json (bench.py, C1, default steps)
# and gpurun_out/r2_c5_1e11_n<N>.json (tools/c5_sweep.sh: 1e11 generated tries of C5 in total).  Every command is bounded.
mkdir -p gpurun_out
N=$1
if [ "$N" = "1" ]; then
  timeout 300 python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/r2_scale_c1_n1.json 2> gpurun_out/r2_scale_c1_n1.err < /dev/null
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2_scale_c1_n$N.json 2> gpurun_out/r2_scale_c1_n$N.err < /dev/null
fi
tail -c 300 gpurun_out/r2_scale_c1_n$N.err
python - <<PY
import json
for line in open("gpurun_out/r2_scale_c1_n$N.json"):
    try: d = json.loads(line)
    except Exception: continue
    print("C1 N=$N: %.4g generated ev/s, %.2f ms/step, e2e %.4g, collective %s ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("allreduce_ms_per_step")))
PY
timeout 600 bash tools/c5_sweep.sh $N < /dev/null
