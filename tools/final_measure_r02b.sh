#!/bin/sh
# Second measurement pass of round 2 (after the 3x3 weight gradient / fp16 3x3 / weight chain-rule kernels): tests, smoke, the bench lines
# whose code paths changed.  Outputs in gpurun_out/r02b_*.
O=gpurun_out
timeout 300 python -m pytest tests -q -m gpu --durations=3 > $O/r02b_pytest.log 2>&1; tail -2 $O/r02b_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02b_smoke.log 2>&1; tail -1 $O/r02b_smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > $O/r02b_bench_R.json 2> $O/r02b_bench_R.err
timeout 150 python bench.py --config T --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_T.json 2> $O/r02b_bench_T.err
timeout 150 python bench.py --config pti --steps 10 --warmup 3 > $O/r02b_bench_pti.json 2>/dev/null
timeout 150 python bench.py --config pti --generator T --steps 10 --warmup 3 > $O/r02b_bench_pti_T.json 2>/dev/null
for f in R T pti pti_T; do python - $O/r02b_bench_$f.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d.get('roofline') or {}
    print(sys.argv[1].split('bench_')[1], round(d['value'], 3), d['unit'], 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 3),
          'frac', r.get('frac'), 'fp32', r.get('fp32_pipe_frac'), 'flrelu ms', r.get('ms_per_step'), 'conv', (d.get('conv') or {}).get('ms_per_step'))
except Exception as e:
    print(sys.argv[1], 'unparsed', e)
PY
done
