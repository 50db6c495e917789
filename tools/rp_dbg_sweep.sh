#!/bin/bash
# diagnostic: time the row-patch layers with pipeline stages disabled (results are wrong by design)
for d in 0 1 2 4 8 3 12 14; do
  SRT_RP_DBG=$d timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('/tmp/b.json'))
pl=d['roofline']['per_layer']
print('dbg=$d', {k: round(pl[k]['ms'],3) for k in ('down2','down3','up4','up5')})
PY
done
