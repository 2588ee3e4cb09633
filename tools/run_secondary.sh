# GPU-box script: ncu captures of the dominant launch of each non-headline configuration.
# Reports are summarised on the box (tools/ncu_summary.py) and deleted: gpurun_out/ is capped at 64 MiB.
set -x
NCU="ncu --set full --clock-control none -f"
cap() {  # name, kernel regex, launches to skip, script argument
  timeout 300 $NCU -k regex:$2 --launch-skip $3 -c 1 -o /tmp/prof_$1 python tools/secondary_once.py $4 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep > gpurun_out/r01_$1_ncu_summary.txt 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page details 2>/dev/null | grep -A3 -i "OPT \|Est. Speedup" | head -60 > gpurun_out/r01_$1_ncu_hints.txt
  rm -f /tmp/prof_$1.ncu-rep
}
cap fwd3d fwd3d 0 3d
cap inv3d inv3d 2 3d
cap matfwd mat_fwd_fused 0 mat
cap matinv mat_inv_fast 11 mat
cap axis1d_fwd axis1d_fast 0 1d
cap axis1d_inv axis1d_inv 9 1d
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_secondary_r1.csv python tools/secondary_once.py > /dev/null 2>&1
du -sh gpurun_out
