"""Per-CUDA-source-line profile of one kernel launch from an .ncu-rep captured with --set full --import-source on (-lineinfo build):
executed warp-instructions and stall samples (~ time) per source line, top N lines, plus totals per file and per line RANGE
(phases of the fused layer kernels).   python tools/ncu_lines.py rep.ncu-rep launch_index [top_n] [ranges file:lo-hi,...]"""
import csv
import io
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ranges = sys.argv[4].split(",") if len(sys.argv) > 4 else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, kern, hdr = [], "?", "?", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        kern = r[1]
    elif r[0] == "Line No":
        hdr = r
        ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].strip().isdigit():
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[ii]), int(r[isamp])))
        except ValueError:
            pass
tot_i, tot_s = sum(l[3] for l in lines), sum(l[4] for l in lines)
print(f"==== {kern[:90]} | warp-inst {tot_i} | samples {tot_s}")
per_file = {}
for f, ln, src, n, s in lines:
    a = per_file.setdefault(f, [0, 0]); a[0] += n; a[1] += s
for f, (n, s) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  file {f:18s} inst {n / tot_i:6.3f} time {s / max(tot_s, 1):6.3f}")
for spec in ranges:
    f, lohi = spec.split(":")
    lo, hi = [int(v) for v in lohi.split("-")]
    n = sum(l[3] for l in lines if l[0] == f and lo <= l[1] <= hi)
    s = sum(l[4] for l in lines if l[0] == f and lo <= l[1] <= hi)
    print(f"  range {spec:28s} inst {n / tot_i:6.3f} time {s / max(tot_s, 1):6.3f}")
print(f"  top {top_n} lines by executed warp-instructions")
for f, ln, src, n, s in sorted(lines, key=lambda l: -l[3])[:top_n]:
    print(f"  {f}:{ln:<5d} inst {n / tot_i:6.3f} time {s / max(tot_s, 1):6.3f} | {src[:110]}")
