# run on the GPU box: DRAM bytes per kernel with ncu (cache state preserved between kernels)
cd tools/l2_hint_probe
i=0
for cfg in "32 512 0 0" "32 512 0 1" "32 512 1 1" "64 512 0 1" "64 512 1 1" "64 128 0 1" "96 512 0 1"; do
  i=$((i+1))
  ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --csv ./probe $cfg 3 > ../../gpurun_out/l2probe_$i.csv 2>/dev/null
  echo "cfg $cfg" >> ../../gpurun_out/l2probe_$i.csv
done
