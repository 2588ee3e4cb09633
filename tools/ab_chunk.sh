# A/B of the chunk size (images per chunk, chunks alternate between two streams) on the headline workload
for cfg in "32 2" "40 2" "44 2" "48 2" "24 2" "22 2" "16 2" "0 2" "32 2"; do
  set -- $cfg
  WTB200_CHUNK=$1 WTB200_STREAMS=$2 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --no-incumbent 2>&1 | tail -1 > /tmp/ab.json
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open('/tmp/ab.json'))
print("chunk", sys.argv[1], "streams", sys.argv[2], "ms", round(d["ms_per_step"], 4), "median", round(d["roofline"]["median_step_ms"], 4), "step_frac", round(d["roofline"]["step_frac"], 4), "clk", d["clocks"]["sm_mhz"])
PY
done
