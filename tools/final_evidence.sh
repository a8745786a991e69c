set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2f_tests.log 2>&1; tail -3 gpurun_out/r2f_tests.log
python bench.py > gpurun_out/r2f_bench_c1.json 2> gpurun_out/r2f_bench_c1.err
for c in c2 c3 c4 c5; do python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_$c.json 2> gpurun_out/r2f_bench_$c.err; done
tools/ncu_capture.sh r2f
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_launch_list.log 2>&1
for f in gpurun_out/r2f_bench_c*.json; do python tools/bench_line.py $f < $f; done
