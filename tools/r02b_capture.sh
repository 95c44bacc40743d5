#!/bin/bash
# Late round-2 evidence (single-image latency work): the bench line, the launch list of ONE single-image 1080p call and
# an ncu --set full capture of the padded k_tail; every ncu pass follows the same command exiting 0 without ncu.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02b_bench_1080p.json 2> gpurun_out/r02b_bench_1080p.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference_1080p.json 2> /dev/null
for s in 1920x1080 3840x2160 640x480; do python tools/latency_breakdown.py $s; python tools/latency_breakdown.py $s imageproc; done > gpurun_out/r02b_latency.txt 2>&1
python tools/latency.py 1920x1080 > gpurun_out/r02b_latency_1080p.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_single_1080p.csv python tools/latency.py 1920x1080 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tail -c 1 -o gpurun_out/r02b_tail -f python tools/latency.py 1920x1080 > gpurun_out/r02b_ncu_tail.log 2>&1
ncu -i gpurun_out/r02b_tail.ncu-rep --page raw --csv > gpurun_out/r02b_tail_raw.csv 2>/dev/null
rm -f gpurun_out/r02b_*.ncu-rep
ls -la gpurun_out/r02b_*
