#!/bin/bash
# developer helper: e2e (host buffer) throughput vs host batch size
for b in "$@"; do
  MRCZIP_HOST_BATCH_CHUNKS=$b python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/e_$b.json
  python - "$b" <<'PY'
import sys, json
b = sys.argv[1]
d = json.loads(open(f"/tmp/e_{b}.json").read())
print(b, round(d["value"], 1), d["e2e"])
PY
done
