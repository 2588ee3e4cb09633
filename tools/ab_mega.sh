# A/B of the persistent all-levels kernel with the scratch-slot ring (headline workload)
run() {  # env assignments...
  env "$@" timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-incumbent 2>&1 | tail -1 > /tmp/ab.json
  python - "$*" <<'PY'
import json, sys
try:
    d = json.load(open('/tmp/ab.json'))
    print(sys.argv[1], "| ms", round(d["ms_per_step"], 4), "step_frac", round(d["roofline"]["step_frac"], 4), "err", d["parity"]["max_rel_err_vs_oracle"], "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[1], "| FAILED", open('/tmp/ab.json').read()[-300:])
PY
}
run WTB200_MEGA=0
run WTB200_MEGA=1 WTB200_MEGA_NOHINTS=1
for ring in 2 3 4; do for seg in 256 128 64; do
  run WTB200_MEGA=1 WTB200_MEGA_NOHINTS=1 WTB200_MEGA_RING=$ring WTB200_MEGA_SEG=$seg
done; done
run WTB200_MEGA=1 WTB200_MEGA_RING=2 WTB200_MEGA_SEG=128
run WTB200_MEGA=1 WTB200_MEGA_RING=3 WTB200_MEGA_SEG=128
