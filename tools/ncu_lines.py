"""Summarise an ncu report per CUDA source line (warp instructions executed, stall samples).
usage: python tools/ncu_lines.py report.ncu-rep [top_n] [kernel_substring]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = None
hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])  # inst, samples, thread_inst, text
seen_fn = set()
per_file = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or r[0] == "" or want and want not in cur_fn:
        continue
    if cur_fn not in seen_fn and len(seen_fn) >= 1 and cur_fn not in seen_fn:
        pass
    seen_fn.add(cur_fn)
    try:
        n = int(r[ci["Instructions Executed"]])
        s = int(r[ci["# Samples"]])
        ti = int(r[ci["Thread Instructions Executed"]])
    except (ValueError, KeyError):
        continue
    key = (cur_fn[:40], cur_file.split("/")[-1], int(r[0]))
    a = agg[key]
    a[0] += n; a[1] += s; a[2] += ti; a[3] = r[1].strip()[:90]
    per_file[(cur_fn[:40], cur_file.split("/")[-1])] += n
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print("total warp inst", tot, "samples", tots)
for k, v in per_file.most_common(12):
    print(f"  {100*v/tot:5.1f}%  {k}")
print("--- top lines by instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/max(tots,1):5.1f}% stall  thr/inst {a[2]/max(a[0],1):4.1f}  {k[1]}:{k[2]}  {a[3]}")
print("--- top lines by stall samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn // 2]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/max(tots,1):5.1f}% stall  thr/inst {a[2]/max(a[0],1):4.1f}  {k[1]}:{k[2]}  {a[3]}")
