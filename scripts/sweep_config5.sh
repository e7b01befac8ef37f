#!/bin/bash
# Config 5: pool-size sweep, entropy round + FI round (B = 10,000 pre-filter, greedy k = 100) + MC-entropy
# round, on the GPUs of this box.  Usage: scripts/sweep_config5.sh <n_gpus> <out.jsonl> [pool sizes per GPU ...]
N=${1:-1}; OUT=${2:-gpurun_out/sweep.jsonl}; shift 2
POOLS=${@:-"10000 100000 1000000 10000000"}
: > "$OUT"
for P in $POOLS; do
  PER=$(( P / N ))
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus "$N" --steps 3 --warmup 3 --no-cpu --pool "$PER" --mc-T 0 2>>"$OUT.err" | grep '^{' >> "$OUT"
  else
    python bench.py --steps 3 --warmup 3 --no-cpu --pool "$PER" --mc-T 0 2>>"$OUT.err" | grep '^{' >> "$OUT"
  fi
done
python - "$OUT" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    fi = d.get('fi_round') or {}
    print('gpus %d pool/gpu %8d: %.3g samples/s resident, %.3g e2e, %.2f ms/round; FI round %.1f ms (greedy %.1f ms)' % (
        d['n_gpus'], d['config']['pool_per_gpu'], d['value'], d['e2e']['value'], d['ms_per_step'],
        fi.get('ms_per_round', float('nan')), (fi.get('stage_ms') or {}).get('greedy', float('nan'))))
PY
