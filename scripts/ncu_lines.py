"""Rank CUDA source lines of one captured launch by stall samples: ncu_lines.py report.ncu-rep [launch_index] [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
idx = sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", idx,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows, cur, hdr = [], None, None
for x in csv.reader(out.splitlines()):
    if len(x) == 2 and x[0] == "File Path":
        cur = x[1].split("/")[-1]
    elif len(x) > 8 and x[0] == "Line No":
        hdr = x
    elif len(x) > 8 and x[0].isdigit() and x[2] == "-":
        rows.append((cur, int(x[0]), x))
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
ts = sum(int(x[si]) for _, _, x in rows)
ti = sum(int(x[ii]) for _, _, x in rows)
print("samples", ts, "warp instructions", ti)
for f, l, x in sorted(rows, key=lambda t: -int(t[2][si]))[:top]:
    st = collections.Counter({hdr[i][6:]: int(x[i] or 0) for i in stall}).most_common(2)
    print(f"{int(x[si]) / ts * 100:5.1f}% {int(x[ii]):>9} {f}:{l:<4} {x[1].strip()[:84]}  {st}")
