#!/bin/bash
# One `ncu --set full` capture of a 4 M-try batch of the loop (16 kernels of C1), reduced ON THE BOX to the text
# summaries we keep (the .ncu-rep itself is too large to travel back):  tools/ncu_capture.sh TAG [config]
# Writes gpurun_out/TAG_ncu_loop.txt (tools/ncu_summary.py) and per-line counts of the largest kernels.
TAG=${1:-rX}
CFG=${2:-c1}
REP=/tmp/${TAG}_ncu.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:"k_generate|k_regen|k_arm|seg_|k_radw|k_finish" --launch-skip 60 --launch-count 20 -f -o ${REP%.ncu-rep} \
  python bench.py --config $CFG --tries 4194304 --batch 4194304 --steps 1 --warmup 3 --no-cpu-baseline \
  > gpurun_out/${TAG}_ncu.log 2>&1 < /dev/null
python tools/ncu_summary.py $REP > gpurun_out/${TAG}_ncu_loop.txt 2>&1 < /dev/null
for k in "k_arm<1, 6>:k_arm16" "k_arm<0, 6>:k_arm06" "k_generate:k_generate" "k_regen:k_regen" "k_finish:k_finish" "k_arm<1, 0>:k_arm10"; do
  re="${k%%:*}"; nm="${k##*:}"
  python tools/ncu_lines.py $REP "$re" 0 60 > gpurun_out/${TAG}_lines_${nm}.txt 2>&1 < /dev/null
done
ls -la $REP
