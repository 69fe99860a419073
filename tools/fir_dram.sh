#!/bin/bash
# developer: DRAM bytes and duration of the persistent FIR kernel (config 3) for a few queue settings
for cfg in "8 4 8" "12 5 10" "16 5 10"; do
  set -- $cfg
  echo "== G=$1 LM=$2 LI=$3"
  TFX_FIR_G=$1 TFX_FIR_LM=$2 TFX_FIR_LI=$3 OS_REPS=2 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fir16k_kernel --launch-skip 1 -c 1 python tools/fir_once.py 2>&1 | grep -E "dram__|gpu__time|lts__"
done
