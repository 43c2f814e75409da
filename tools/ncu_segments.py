"""Summarise an .ncu-rep: per kernel, duration / issue utilisation / stall mix, and the share of executed
warp-instructions between consecutive BAR.SYNC instructions (= the phases of the fused layer kernels)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
names = []
for r in rows[2:]:
    d = dict(zip(h, r)); names.append(d['Kernel Name'])
    print('====', d['Kernel Name'][:60])
    for k in keys:
        if k in d: print(f"   {k.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio','')} = {d[k]}")
for idx, name in enumerate(names):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; ia = hdr.index('Source'); ii = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
    seg = []; cur = [0, 0, 0, {}]; tot = 0
    for r in rows[2:]:
        try: n = int(r[ii]); s = int(r[isamp])
        except Exception: continue
        toks = r[ia].strip().split(); op = toks[0] if toks else ''
        if op.startswith('@') and len(toks) > 1: op = toks[1]
        cur[0] += n; cur[1] += s; cur[2] += 1
        key = op.split('.')[0]; cur[3][key] = cur[3].get(key, 0) + n; tot += n
        if 'BAR.SYNC' in r[ia] or 'EXIT' in r[ia]:
            seg.append(cur); cur = [0, 0, 0, {}]
    seg.append(cur); ts = sum(s[1] for s in seg) or 1
    print('---- phases of', name[:50], 'total warp-inst', tot)
    for i, s in enumerate(seg):
        if s[0] / max(tot, 1) < 0.004: continue
        top = sorted(s[3].items(), key=lambda kv: -kv[1])[:8]
        print(f"  seg{i:2d} inst={s[0]/tot:6.3f} samples={s[1]/ts:6.3f} sass={s[2]:4d} ", ' '.join(f"{k}:{v/max(s[0],1):.2f}" for k, v in top))
