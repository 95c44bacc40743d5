#!/bin/bash
# ncu --set full (with source) of the 11-tap marching blur on the octaves of one 32-image 1080p group (development aid)
S="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra"
SB200_GRAPHS=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_blur_march<\(int\)1,' --launch-skip 12 -c 4 -o gpurun_out/s4_b1 -f $S > gpurun_out/s4_ncu.log 2>&1
ncu -i gpurun_out/s4_b1.ncu-rep --page raw --csv > gpurun_out/s4_b1_raw.csv 2>/dev/null
ncu -i gpurun_out/s4_b1.ncu-rep --page source --csv > gpurun_out/s4_b1_src.csv 2>/dev/null
rm -f gpurun_out/s4_b1.ncu-rep
ls -la gpurun_out/s4_*
