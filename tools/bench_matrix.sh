#!/bin/bash
# developer helper: bench line per (distribution, bits) for the results table
for kb in "$@"; do
  k=${kb:0:1}; b=${kb:1}
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-matrix --kind $k --bits $b 2>/dev/null > /tmp/m_$kb.json
  python - "$kb" <<'PY'
import sys, json
kb = sys.argv[1]
d = json.loads(open(f"/tmp/m_{kb}.json").read())
print(kb, {k: round(d[k], 4) for k in ["value", "ms_per_step", "compress_GBs", "decompress_GBs", "ratio"]},
      {k: round(v, 2) for k, v in d["stage_ms"].items() if v > 0.3}, d["encode_stats"], d["decode_stats"])
PY
done
