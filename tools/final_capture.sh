set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 3000 gpurun_out/final_bench.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_ref.json 2>&1; tail -1 gpurun_out/final_ref.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_ncu_list.log 2>&1; tail -2 gpurun_out/final_ncu_list.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k 'regex:k_upsample|k_blur_march|k_extrema_tma|k_tail|k_refine|k_orient|k_descriptor' -c 30 -o gpurun_out/final_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/final_ncu_full.log 2>&1; tail -2 gpurun_out/final_ncu_full.log | cut -c1-200
ncu -i gpurun_out/final_full.ncu-rep --page raw --csv > gpurun_out/final_full_raw.csv 2>/dev/null; wc -l gpurun_out/final_full_raw.csv; ls -la gpurun_out/
