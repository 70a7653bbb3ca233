#!/bin/bash
# developer helper: bench at several batch sizes
for b in "$@"; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --batch-chunks $b 2>/dev/null > /tmp/b_$b.json
  python - "$b" <<'PY'
import sys, json
b = sys.argv[1]
d = json.loads(open(f"/tmp/b_{b}.json").read())
print(b, {k: round(d[k], 2) for k in ["value", "ms_per_step", "compress_GBs", "decompress_GBs"]},
      {k: round(v, 2) for k, v in d["stage_ms"].items() if v > 0.3})
PY
done
