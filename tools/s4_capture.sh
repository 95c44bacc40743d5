#!/bin/bash
# ncu --set full (with source) of one marching blur (tap set $1, default 1) on the first four octaves of one 32-image
# 1080p group (development aid); output gpurun_out/s4_b$1_{raw,src}.csv
L=${1:-1}
S="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra"
SB200_GRAPHS=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_blur_march<\(int\)$L," --launch-skip 12 -c 4 -o gpurun_out/s4_b$L -f $S > gpurun_out/s4_ncu.log 2>&1
ncu -i gpurun_out/s4_b$L.ncu-rep --page raw --csv > gpurun_out/s4_b${L}_raw.csv 2>/dev/null
ncu -i gpurun_out/s4_b$L.ncu-rep --page source --csv > gpurun_out/s4_b${L}_src.csv 2>/dev/null
rm -f gpurun_out/s4_b$L.ncu-rep
ls -la gpurun_out/s4_*
