# Profiling pass (run on the GPU box through gpurun): plain bench first, then the ncu launch list of the
# same command, then one --set full capture over one forward pass + the DTW kernel.
set -x
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader,nounits -i 0 2>&1 | head -3
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log | cut -c1-200
# one whole forward (31 launches) of the first timed step: warm-up 1 = launches 0..30, capture 31..61
ncu --set full --clock-control none --import-source on -s 31 -c 31 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-align > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"dtw_pipeline" -s 1 -c 1 -o gpurun_out/${TAG}_dtw python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu3.log | cut -c1-200
ls -la gpurun_out | tail -8
