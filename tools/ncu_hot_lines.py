"""Top source lines by warp-stall samples from an .ncu-rep captured with --import-source on (kernel compiled -lineinfo).

    python tools/ncu_hot_lines.py gpurun_out/foo.ncu-rep [N]
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
js = hdr.index("Warp Stall Sampling (All Samples)")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) != len(hdr):
        if r and r[0] == "Line No":
            break           # next kernel instance
        continue
    a = agg.setdefault(r[0], [r[1].strip(), 0, 0])
    a[1] += int(r[js]) if r[js].isdigit() else 0
    a[2] += 1
tot = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot}")
for ln, (src, n, k) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ln:>5s} {100.0 * n / tot:5.1f}%  {src[:130]}")
