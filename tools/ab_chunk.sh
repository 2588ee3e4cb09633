# A/B of the chunked / scratch-reusing forward path on the headline workload
for cfg in "0 2" "1 2" "2 2" "3 2" "4 2" "8 2" "1 1" "2 1" "4 1"; do
  set -- $cfg
  WTB200_CHUNK=$1 WTB200_STREAMS=$2 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-incumbent 2>&1 | tail -1 > /tmp/ab.json
  python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open('/tmp/ab.json'))
print("chunk", sys.argv[1], "streams", sys.argv[2], "ms", round(d["ms_per_step"], 4), "step_frac", round(d["roofline"]["step_frac"], 4), "err", d["parity"]["max_rel_err_vs_oracle"], "launches", d["gpu_launches"])
PY
done
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-incumbent 2>&1 | tail -1 | python -c "
import json,sys; d=json.load(sys.stdin); print('default', round(d['ms_per_step'],4), round(d['roofline']['step_frac'],4), d['parity'])"
