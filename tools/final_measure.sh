#!/bin/sh
# Round-end measurement pass on one B200: tests, reference-plugin comparison, every bench line, ncu captures.  Outputs in gpurun_out/r02_final_*.
O=gpurun_out
python -m pytest tests -q -m gpu --durations=5 > $O/r02_final_pytest.log 2>&1; tail -3 $O/r02_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1; tail -1 $O/r02_final_smoke.log
python tools/check_vs_reference_cuda.py --out $O/r02_final_vs_reference_plugin.md > $O/r02_final_vsref.log 2>&1; tail -1 $O/r02_final_vsref.log
python bench.py --steps 20 --warmup 5 > $O/r02_final_bench_R.json 2> $O/r02_final_bench_R.err
python bench.py --config T --steps 10 --warmup 3 > $O/r02_final_bench_T.json 2> $O/r02_final_bench_T.err
python bench.py --math fp32x3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_final_bench_R_fp32x3.json 2>/dev/null
python bench.py --config restyle --steps 3 --warmup 3 > $O/r02_final_bench_restyle.json 2>/dev/null
python bench.py --config pti --steps 20 --warmup 3 > $O/r02_final_bench_pti.json 2>/dev/null
python bench.py --impl reference --steps 1 --warmup 0 > $O/r02_final_bench_reference.json 2>/dev/null
for f in R T R_fp32x3 restyle pti reference; do python - $O/r02_final_bench_$f.json <<'PY'
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
python tools/prof_ops.py > $O/r02_final_streaming_ops.txt 2>&1; tail -6 $O/r02_final_streaming_ops.txt
[ -n "$SKIP_NCU" ] && exit 0
cap() { ncu --set full --import-source on --clock-control none -k kernel -s $3 -c 1 -o $O/r02_final_$1 python tools/prof_flrelu.py $2 2 bwd > /dev/null 2>&1; }
cap L11_fwd L11 3; cap L10_fwd L10 3; cap L12_fwd L12 3; cap L11_write L11 4; cap L11_bwd L11 5; cap L10_bwd L10 5
ls $O/r02_final_*.ncu-rep
