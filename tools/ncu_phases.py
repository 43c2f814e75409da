"""Per-phase breakdown of a fused layer kernel from an .ncu-rep (--set full --import-source on): the SASS between consecutive
BAR.SYNC instructions is one phase; prints its share of executed warp-instructions, of the stall samples (~ time), the
top opcodes and the top stall reasons.   python tools/ncu_phases.py rep.ncu-rep [launch_index ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
launches = [int(v) for v in sys.argv[2:]] or [0]
for idx in launches:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = rows[1]; ia = hdr.index('Source'); ii = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    segs = []; cur = dict(n=0, s=0, k=0, ops={}, st={}); tot = 0; tots = 0
    for r in rows[2:]:
        try: n = int(r[ii]); s = int(r[isamp])
        except Exception: continue
        toks = r[ia].strip().split(); op = toks[0] if toks else ''
        if op.startswith('@') and len(toks) > 1: op = toks[1]
        cur['n'] += n; cur['s'] += s; cur['k'] += 1; tot += n; tots += s
        key = op.split('.')[0]; cur['ops'][key] = cur['ops'].get(key, 0) + n
        for i, h in stall_cols:
            try: cur['st'][h] = cur['st'].get(h, 0) + int(r[i])
            except Exception: pass
        if 'BAR.SYNC' in r[ia] or 'EXIT' in r[ia]:
            segs.append(cur); cur = dict(n=0, s=0, k=0, ops={}, st={})
    segs.append(cur)
    print('==== launch', idx, name[:70], '| warp-inst', tot, '| samples', tots)
    for i, s in enumerate(segs):
        if s['s'] / max(tots, 1) < 0.005 and s['n'] / max(tot, 1) < 0.005: continue
        top = sorted(s['ops'].items(), key=lambda kv: -kv[1])[:7]
        st = sorted(s['st'].items(), key=lambda kv: -kv[1])[:5]
        print(f"seg{i:2d} inst={s['n']/tot:6.3f} time={s['s']/tots:6.3f} sass={s['k']:4d} | " + ' '.join(f"{k}:{v/max(s['n'],1):.2f}" for k, v in top))
        print("        stalls: " + ' '.join(f"{k[6:]}:{v/max(s['s'],1):.2f}" for k, v in st))
