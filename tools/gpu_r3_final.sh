#!/bin/bash
# end of round 2, session 4: whole GPU suite, smoke(), what the driver runs (both arms, N = 1) + the other configurations, fp32_tc lines
mkdir -p gpurun_out/r3
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  *\[\|^E    *[0-9-]" | tail -30 > gpurun_out/r3/final_test_all_gpu.txt; tail -3 gpurun_out/r3/final_test_all_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r3/final_bench_reference.json 2> gpurun_out/r3/final_bench_reference.err
( time python bench.py ) > gpurun_out/r3/final_bench_c2.json 2> gpurun_out/r3/final_bench_c2.err
tail -4 gpurun_out/r3/final_bench_c2.err
for w in c1 c5; do python bench.py --workload $w --no-cpu-baseline --no-eager-baseline --no-extras --steps 10 > gpurun_out/r3/final_bench_$w.json 2> gpurun_out/r3/final_bench_$w.err; done
for w in c2 c3 c4 c5; do python bench.py --workload $w --mode infer --no-cpu-baseline --steps 10 > gpurun_out/r3/final_infer_$w.json 2> gpurun_out/r3/final_infer_$w.err; done
for w in c3 c4; do python bench.py --workload $w --precision fp32_tc --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-extras > gpurun_out/r3/final_bench_${w}_fp32_tc.json 2> gpurun_out/r3/final_bench_${w}_fp32_tc.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3/final_*.json')):
    try:
        l=json.load(open(f))
        print(f.split('/')[-1], l.get('impl','ours'), l['metric'], round(l['value']), 'e2e', round(l['e2e']['value']), 'ms', round(l['ms_per_step'],2))
        for k in ('infer','other_workloads','cuda_eager_baseline','cpu_baseline'):
            if k in l: print('    ',k, json.dumps(l[k])[:600])
    except Exception as e: print(f, 'ERR', e)
PY
