#!/bin/bash
# usage: tools/ab_env.sh VAR v1 v2 ... -- runs the forward bench with VAR=v and prints ms/step + step_frac
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 > /tmp/ab.json
  python - "$var" "$v" <<'PY'
import json, sys
d = json.load(open('/tmp/ab.json'))
print(sys.argv[1], sys.argv[2], round(d["ms_per_step"], 4), round(d["roofline"]["step_frac"], 4), d["parity"]["max_rel_err_vs_oracle"], d["gpu_launches"])
PY
done
