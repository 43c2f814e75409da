#!/bin/bash
# what the driver runs at round end (both arms, N = 1) + the other configurations for README / profiles
mkdir -p gpurun_out/r2
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2/final_bench_reference.json 2> gpurun_out/r2/final_bench_reference.err
( time python bench.py ) > gpurun_out/r2/final_bench_c2.json 2> gpurun_out/r2/final_bench_c2.err
tail -4 gpurun_out/r2/final_bench_c2.err
for w in c1 c5; do python bench.py --workload $w --no-cpu-baseline --no-eager-baseline --no-extras --steps 10 > gpurun_out/r2/final_bench_$w.json 2> gpurun_out/r2/final_bench_$w.err; done
for w in c2 c3 c4 c5; do python bench.py --workload $w --mode infer --no-cpu-baseline --steps 10 > gpurun_out/r2/final_infer_$w.json 2> gpurun_out/r2/final_infer_$w.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2/final_*.json')):
    try:
        l=json.load(open(f))
        print(f.split('/')[-1], l.get('impl','ours'), l['metric'], round(l['value']), 'e2e', round(l['e2e']['value']), 'ms', round(l['ms_per_step'],2))
        for k in ('infer','other_workloads','cuda_eager_baseline','cpu_baseline'):
            if k in l: print('    ',k, json.dumps(l[k])[:400])
    except Exception as e: print(f, 'ERR', e)
PY
