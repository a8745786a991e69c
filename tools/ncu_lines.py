#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel of an .ncu-rep (cuda,sass view):
   tools/ncu_lines.py file.ncu-rep KERNEL_REGEX LAUNCH_SKIP [TOP]"""
import csv, subprocess, sys, collections
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
fname = None
hdr = None
per_line = collections.Counter()
per_file = collections.Counter()
samples = collections.Counter()
text = {}
cur = None
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Name":
        fname = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr is None or len(row) < len(hdr):
        continue
    if row[0] != "":
        cur = (fname, int(row[0]))
        text[cur] = row[1].strip()[:90]
        continue
    try:
        n = int(row[i_inst]); s = int(row[i_samp])
    except ValueError:
        continue
    per_line[cur] += n
    per_file[fname] += n
    samples[cur] += s
tot = sum(per_line.values())
tots = sum(samples.values())
print("total warp instructions", tot, "samples", tots)
for f, n in per_file.most_common():
    print(f"  {str(f):24s} {100.0 * n / tot:5.1f}%")
print("top lines by instructions (inst%, samples%)")
for k, n in per_line.most_common(top):
    print(f"  {100.0 * n / tot:5.1f}% {100.0 * samples[k] / max(tots, 1):5.1f}%  {k[0]}:{k[1]}  {text.get(k, '')}")
