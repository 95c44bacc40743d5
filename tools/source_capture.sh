#!/bin/bash
# SASS-level hot spots (opcode mix, stall reasons, hottest lines) of the three heaviest kernels, from bench.py itself
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k 'regex:k_descriptor|k_orient' -c 2 -o gpurun_out/src_kp -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/src_kp.log 2>&1
ncu -i gpurun_out/src_kp.ncu-rep --page source --csv > gpurun_out/src_kp.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_blur_march --launch-skip 5 -c 1 -o gpurun_out/src_b5 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/src_b5.log 2>&1
ncu -i gpurun_out/src_b5.ncu-rep --page source --csv > gpurun_out/src_b5.csv 2>/dev/null
for k in 0 2; do KSEL=$k python tools/ncu_source.py gpurun_out/src_kp.csv 40; echo; done > gpurun_out/src_kp.txt 2>&1
KSEL=0 python tools/ncu_source.py gpurun_out/src_b5.csv 40 > gpurun_out/src_b5.txt 2>&1
rm -f gpurun_out/*.ncu-rep
head -3 gpurun_out/src_kp.txt gpurun_out/src_b5.txt; tail -3 gpurun_out/src_b5.log
