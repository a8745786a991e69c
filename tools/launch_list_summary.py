"""Per-stage and per-kernel sums of an ncu launch list (gpu__time_duration.sum, --csv) of bench.py, next to the
stage_ms shares of a bench line:  python tools/launch_list_summary.py profiles/r2f_launch_list.csv profiles/r2f_bench_c1.json
A seg_N kernel (compiled stretch) belongs to the arm of the k_arm<W, *> launch in front of it."""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt, stage, arm = collections.OrderedDict(), collections.Counter(), collections.Counter(), None
for r in rows:
    k, ns = r[ki], float(r[vi])
    short = k.replace("simc::", "").replace("strict::", "").split("(")[0].replace("void ", "")
    if "fp64_peak" in k:
        st = "(fp64 peak microbenchmark)"
    elif "k_generate" in k or "k_regen" in k:
        st = "k_generate"
    elif "k_arm<1" in k:
        arm = st = "k_arm<hadron>"
    elif "k_arm<0" in k:
        arm = st = "k_arm<electron>"
    elif k.startswith("seg_"):
        st = arm
        short = f"{short} ({'P' if arm == 'k_arm<hadron>' else 'E'})"
    elif "k_radw" in k or "k_finish" in k:
        st = "k_finish"
    else:
        st = "(other)"
    tot[short] = tot.get(short, 0) + ns
    cnt[short] += 1
    stage[st] += ns
loop = sum(v for k, v in stage.items() if not k.startswith("("))
print(f"{sys.argv[1]}: {len(rows)} launches; loop kernels {loop / 1e6:.1f} ms")
for k, v in stage.items():
    print(f"  stage {k:30s} {v / 1e6:9.2f} ms  {('%.1f %%' % (100 * v / loop)) if not k.startswith('(') else ''}")
if len(sys.argv) > 2:
    d = json.load(open(sys.argv[2]))["roofline"]["stage_ms"]
    s = sum(d.values())
    print(f"bench stage_ms shares ({sys.argv[2]}):", {k: round(100 * v / s, 1) for k, v in d.items()})
print("per kernel:")
for k, v in tot.items():
    print(f"  {k:40s} n={cnt[k]:3d}  {v / 1e6:9.3f} ms  avg {v / cnt[k] / 1e6:7.3f} ms")
