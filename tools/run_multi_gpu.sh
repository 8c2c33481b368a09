#!/bin/sh
# Multi-GPU runs of the configs[3] / configs[4] workloads and of the forward bench on N GPUs of one box:
#   gpurun --gpus N -- sh tools/run_multi_gpu.sh N     -> gpurun_out/r02_multi_nN.jsonl (one JSON line per run)
N=${1:-2}
OUT=gpurun_out/r02_multi_n$N.jsonl
: > $OUT
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" 2>gpurun_out/r02_multi_n$N.err | tail -1 >> $OUT
}
run --config pti --steps 20 --warmup 3
run --config restyle --steps 3 --warmup 3
[ "$N" = "2" ] && run --steps 5 --warmup 3 --no-cpu-baseline
python - "$OUT" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    try:
        d = json.loads(line)
    except Exception:
        print('unparsed:', line[:200]); continue
    print(d['metric'][:60], '| n_gpus', d['n_gpus'], '| value', round(d['value'], 2), d['unit'], '| ms/step', round(d['ms_per_step'], 2),
          '| e2e', round(d['e2e']['value'], 2), '|', d['config'].get('parallelism', ''))
PY
