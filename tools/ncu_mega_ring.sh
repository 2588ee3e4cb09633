for cfg in "0 256" "3 128" "4 128" "2 128"; do
  set -- $cfg
  WTB200_MEGA=1 WTB200_MEGA_NOHINTS=1 WTB200_MEGA_RING=$1 WTB200_MEGA_SEG=$2 timeout 200 ncu --cache-control none --clock-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:mega --launch-skip 3 -c 1 --csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-incumbent 2>/dev/null | grep -E '^"' | awk -F'","' -v c="ring $1 seg $2" 'NR>1{gsub(/"/,"",$NF); printf "%s | %s %s %s\n", c, $(NF-2), $(NF-1), $NF}'
done
