#!/bin/bash
# Round-2 evidence for profiles/ (run on the GPU box through gpurun; every ncu pass follows the same command exiting 0
# without ncu).  Produces under gpurun_out/: the bench lines of both arms, per-launch CUDA-event profiles, launch lists
# of the 1080p and 640x480 bench workloads and of ONE single-image 1080p call, and --set full captures (source view
# included) of the heaviest kernels.  tools/r02_profiles.py turns them into the files of profiles/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_1080p.json 2> gpurun_out/r02_bench_1080p.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_1080p.json 2> /dev/null
python tools/fine_profile.py 1080p > gpurun_out/r02_fine_1080p.txt 2>&1
python tools/fine_profile.py vga > gpurun_out/r02_fine_vga.txt 2>&1
python tools/fine_profile.py 4k > gpurun_out/r02_fine_4k.txt 2>&1
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
$B > gpurun_out/r02_plain_1080p.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_launches_1080p_b32.csv $B > /dev/null 2>&1
$B --workload vga > gpurun_out/r02_plain_vga.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_launches_vga_b128.csv $B --workload vga > /dev/null 2>&1
python tools/latency.py 1920x1080 > gpurun_out/r02_latency_1080p.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_single_1080p.csv python tools/latency.py 1920x1080 > /dev/null 2>&1
S="python bench.py --steps 1 --warmup 3 --no-cpu --no-extra"
ncu --set full --clock-control none --import-source on -k regex:k_descriptor -c 1 -o gpurun_out/r02_desc -f $S > gpurun_out/r02_ncu_desc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_orient -c 1 -o gpurun_out/r02_orient -f $S > gpurun_out/r02_ncu_orient.log 2>&1
ncu --set full --clock-control none -k 'regex:k_blur_march|k_extrema_tma|k_upsample' -c 14 -o gpurun_out/r02_pyr -f $S > gpurun_out/r02_ncu_pyr.log 2>&1
M="python bench.py --workload match --steps 2 --warmup 3 --no-cpu"
$M > gpurun_out/r02_plain_match.log 2>&1 &&
ncu --set full --clock-control none -k regex:k_match_nn -c 1 -o gpurun_out/r02_match -f $M > gpurun_out/r02_ncu_match.log 2>&1
for n in desc orient pyr match; do
  ncu -i gpurun_out/r02_$n.ncu-rep --page raw --csv > gpurun_out/r02_${n}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/r02_desc.ncu-rep --page source --csv > gpurun_out/r02_desc_src.csv 2>/dev/null
ncu -i gpurun_out/r02_orient.ncu-rep --page source --csv > gpurun_out/r02_orient_src.csv 2>/dev/null
rm -f gpurun_out/r02_*.ncu-rep
ls -la gpurun_out/r02_*
