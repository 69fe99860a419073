#!/bin/bash
# developer sweep of the persistent FIR kernel's queue parameters (open slots G, MAC lag LM, inverse lag LI)
for cfg in "8 3 6" "8 4 8" "10 4 8" "10 4 9" "12 4 9" "12 5 10" "16 5 10" "16 6 12"; do
  set -- $cfg
  echo "== G=$1 LM=$2 LI=$3"
  TFX_FIR_G=$1 TFX_FIR_LM=$2 TFX_FIR_LI=$3 timeout 100 python tools/fir_trace.py 2>&1 | grep -E "span|wait mean|busy"
done
