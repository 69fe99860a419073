#!/bin/bash
# developer A/B: staggered warp halves in the forward (bit 0) / inverse (bit 1) transforms of the FIR kernel
for st in 3 0 1 2; do
  echo "== TFX_FIR_STAGGER=$st"
  TFX_FIR_STAGGER=$st timeout 100 python tools/fir_trace.py 2>&1 | grep -E "span|run mean|busy"
done
