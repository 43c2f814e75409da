"""Like ncu_lines.py, but the top N source lines by STALL SAMPLES (~ time), with the dominant stall reasons of each line.
python tools/ncu_lines_time.py rep.ncu-rep launch_index [top_n]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines, fname, hdr = [], "?", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and not h.endswith("_not_issued")]
    elif hdr and r[0].strip().isdigit():
        try:
            st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:3]
            lines.append((fname, int(r[0]), r[1].strip(), int(r[ii]), int(r[isamp]), st))
        except ValueError:
            pass
tot_i, tot_s = sum(l[3] for l in lines), sum(l[4] for l in lines)
print(f"warp-inst {tot_i} samples {tot_s}")
for f, ln, src, n, s, st in sorted(lines, key=lambda l: -l[4])[:top_n]:
    why = " ".join(f"{h[6:]}={v}" for v, h in st if v)
    print(f"  {f}:{ln:<5d} time {s / max(tot_s, 1):6.3f} inst {n / tot_i:6.3f} | {why} | {src[:90]}")
