# Round-1 profiling pass (run on the GPU box through gpurun): plain bench first, then the ncu launch
# list of the same command, then one --set full capture of the dominant kernels.
set -x
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader,nounits -i 0 2>&1 | head -3
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"gcn_fused_kernel|tconv_window_kernel|tc_gemm_kernel|dtw_wavefront" -s 12 -c 14 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu2.log | cut -c1-200
ls -la gpurun_out | tail -8
