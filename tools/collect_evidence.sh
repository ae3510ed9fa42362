#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): ncu launch list of the bench command and one
# ncu --set full capture of every tcgen05 conv launch of one frame, exported to CSV on the box
# (the .ncu-rep itself is too large to travel back).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none -k regex:conv_tc_kernel -c 21 \
    -o /tmp/conv_tc_full -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/conv_tc_full.ncu-rep --page raw --csv > gpurun_out/conv_tc_full_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/conv_tc_full.ncu-rep
