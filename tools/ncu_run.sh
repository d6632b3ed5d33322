set -x
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -i 0 2>&1 | head -3
python bench.py --steps 2 --warmup 1 --batch 64 --pairs 512 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --batch 64 --pairs 512 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-300
python bench.py --steps 1 --warmup 1 --batch 64 --pairs 512 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|dtw_wavefront|aggregate" -s 20 -c 8 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --batch 64 --pairs 512 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log | cut -c1-300
ls -la gpurun_out
