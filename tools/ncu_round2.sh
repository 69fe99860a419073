#!/bin/bash
# Round-2 ncu captures (one B200, run through gpurun): --set full of every kernel that had none, plus the launch list of bench.py.
# Numbers printed under ncu are evidence of where time goes / DRAM traffic, never bench values.
NCU="ncu --set full --clock-control none --import-source on -c 1"
out=gpurun_out
OS_REPS=2 timeout 500 $NCU -k regex:fir16k_kernel --launch-skip 1 -o $out/r2_fir16k_cfg3 -f python tools/fir_once.py > $out/r2n_fir16k.log 2>&1
OS_WHAT=fir_direct OS_REPS=1 timeout 300 $NCU -k regex:fir_direct_kernel --launch-skip 1 -o $out/r2_fir_direct -f python tools/misc_once.py > $out/r2n_fir_direct.log 2>&1
OS_WHAT=fir_f64 OS_REPS=1 timeout 300 $NCU -k regex:fir_direct_f64_kernel --launch-skip 1 -o $out/r2_fir_f64 -f python tools/misc_once.py > $out/r2n_fir_f64.log 2>&1
OS_WHAT=delay OS_REPS=1 timeout 300 $NCU -k regex:delay_kernel --launch-skip 1 -o $out/r2_delay -f python tools/misc_once.py > $out/r2n_delay.log 2>&1
OS_WHAT=bank_stream OS_REPS=1 timeout 300 $NCU -k regex:bank_stream_kernel --launch-skip 3 -o $out/r2_bank_stream -f python tools/misc_once.py > $out/r2n_bank_stream.log 2>&1
OS_WHAT=mixed6 OS_REPS=1 timeout 300 $NCU -k regex:sos_tile_kernel --launch-skip 3 -o $out/r2_tile_mixed6 -f python tools/misc_once.py > $out/r2n_mixed6.log 2>&1
OS_MODE=sum OS_N=32 OS_C=256 OS_T=2880000 OS_PREC=f32 timeout 400 $NCU -k regex:bank_stack_kernel --launch-skip 3 -o $out/r2_sum32_f32 -f python tools/bank_once.py > $out/r2n_sum32.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > $out/r2n_benchncu.log 2>&1
for w in delay fir_direct fir_f64 bank_stream mixed6; do OS_WHAT=$w timeout 200 python tools/misc_once.py; done > $out/r2n_misc_times.log 2>&1
ls -la $out/r2_*.ncu-rep
# added at the end of round 2: the 32-channel STACK shard after the band dealing, and the time-parallel warm-up kernel
OS_MODE=stack OS_C=32 OS_T=2880000 OS_PREC=auto timeout 400 $NCU -k regex:bank_stack_kernel --launch-skip 1 -o $out/r2_stack32ch_b -f python tools/bank_once.py > $out/r2n_stack32b.log 2>&1
OS_MODE=stack OS_C=32 OS_T=2880000 OS_PREC=auto timeout 400 $NCU -k regex:bank_warm1_kernel --launch-skip 1 -o $out/r2_warm1_32ch_c -f python tools/bank_once.py > $out/r2n_warm1.log 2>&1
