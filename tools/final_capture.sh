#!/bin/bash
# Captures what profiles/ holds, from bench.py itself (run on the GPU box through gpurun).  FULL=1 also runs the GPU
# tests and the reference arm first.
set -x
mkdir -p gpurun_out
if [ -n "$FULL" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_ref.json 2>&1; tail -1 gpurun_out/final_ref.json | cut -c1-600
fi
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_ncu_list.log 2>&1
ncu --set full --clock-control none -k 'regex:k_upsample|k_blur_march|k_extrema_tma' -c 14 -o gpurun_out/final_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/final_ncu_full.log 2>&1
ncu -i gpurun_out/final_full.ncu-rep --page raw --csv > gpurun_out/final_full_raw.csv 2>/dev/null
ncu --set full --clock-control none -k 'regex:k_tail|k_compact|k_refine|k_orient|k_descriptor' -c 5 -o gpurun_out/final_kp -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/final_ncu_kp.log 2>&1
ncu -i gpurun_out/final_kp.ncu-rep --page raw --csv > gpurun_out/final_kp_raw.csv 2>/dev/null
rm -f gpurun_out/final_kp.ncu-rep gpurun_out/final_full.ncu-rep
wc -l gpurun_out/final_full_raw.csv gpurun_out/final_kp_raw.csv
