"""profiles/r2_traffic.json from a tools/ncu_summary.py text file of one batch of the loop:

    python tools/make_traffic.py profiles/r2f_ncu_loop.txt "state described in one line" > profiles/r2_traffic.json

Adds dram__bytes_read.sum + dram__bytes_write.sum (GB in the summary) and gpu__time_duration.sum (ms) of the kernels
of each stage of bench.py's work model.  The capture window may start in the middle of a batch: the launches are
walked cyclically from k_generate to the launch before the next k_generate; a seg_N kernel belongs to the arm of the
k_arm<W, *> launch in front of it.
"""
import ast
import json
import sys


def rows_of(path):
    rows, units = {}, {}
    for line in open(path):
        i = line.find("[")
        if i < 0:
            continue
        try:
            head = line[:i].split()
            rows[head[0] if head[0] != "Kernel" else "Kernel Name"] = ast.literal_eval(line[i:].strip())
            units[head[0]] = head[1] if len(head) > 1 else ""
        except (ValueError, SyntaxError):
            pass
    return rows, units


def main():
    path = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else ""
    rows, units = rows_of(path)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}
    b_rd, b_wr, t_ms = (scale[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    names = rows["Kernel Name"]
    n = len(names)
    start = next(i for i, k in enumerate(names) if "k_generate" in k)
    order = []
    for j in range(n):
        i = (start + j) % n
        if j > 0 and "k_generate" in names[i]:
            break
        order.append(i)
    stage_of, arm = {}, None
    for i in order:
        k = names[i]
        if "k_generate" in k or "k_regen" in k:
            st = "k_generate"
        elif "k_arm<1" in k or "k_calo" in k:
            arm = st = "k_arm<hadron>"
        elif "k_arm<0" in k:
            arm = st = "k_arm<electron>"
        elif k.startswith("seg_"):
            st = arm
        else:
            st = "k_finish"                       # k_radw, k_finish
        stage_of[i] = st
    out = {"source": f"{path}: ncu --set full --clock-control none, bench.py --tries 4194304 --batch 4194304 "
                     f"(tools/ncu_capture.sh{', ' + what if what else ''}); dram__bytes_read.sum + dram__bytes_write.sum, "
                     "all kernels of a stage added (tools/make_traffic.py)",
           "bytes_per_4M_tries": {}, "ms_under_ncu": {}, "kernels": {}}
    rd, wr, ms = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"], rows["gpu__time_duration.sum"]
    seen = set()
    for i in order:
        st = stage_of[i]
        if (st, names[i]) in seen:                # a window longer than one batch repeats its first launches
            continue
        seen.add((st, names[i]))
        out["bytes_per_4M_tries"][st] = out["bytes_per_4M_tries"].get(st, 0.0) + float(rd[i]) * b_rd + float(wr[i]) * b_wr
        out["ms_under_ncu"][st] = out["ms_under_ncu"].get(st, 0.0) + float(ms[i]) * t_ms
        out["kernels"].setdefault(st, []).append(names[i][:40])
    tot = sum(out["ms_under_ncu"].values())
    out["share_under_ncu"] = {k: round(v / tot, 4) for k, v in out["ms_under_ncu"].items()}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
