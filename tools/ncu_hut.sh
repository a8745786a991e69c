#!/bin/bash
# Source-level capture of the hut kernel of the hadron arm (k_arm<1, 6>, first k_arm launch after the skip) of one C1 batch:
#   tools/ncu_hut.sh TAG  ->  gpurun_out/TAG_lines_hut.txt
TAG=${1:-rX}
REP=/tmp/${TAG}_hut.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_arm" --launch-skip 25 --launch-count 1 -f -o ${REP%.ncu-rep} \
  python bench.py --config c1 --tries 4194304 --batch 4194304 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_hut.log 2>&1 < /dev/null
ncu -i $REP --page raw --csv 2>/dev/null | head -3 | cut -c1-400 > gpurun_out/${TAG}_hut_name.txt
python tools/ncu_lines.py $REP "k_arm" 0 90 > gpurun_out/${TAG}_lines_hut.txt 2>&1 < /dev/null
