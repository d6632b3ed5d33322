# Profiling pass (run on the GPU box through gpurun): plain bench first, then the ncu launch list of the
# same command, then one --set full capture over one forward pass + the DTW kernel.
set -x
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader,nounits -i 0 2>&1 | head -3
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log | cut -c1-200
# one whole forward (25 launches) of the second pass: the first pass = launches 0..24, capture 25..49
python tools/prof_seg.py > gpurun_out/${TAG}_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 25 -c 25 -o gpurun_out/${TAG}_full python tools/prof_seg.py > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"dtw_ws" -s 1 -c 1 -o gpurun_out/${TAG}_dtw $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu3.log | cut -c1-200
ls -la gpurun_out | tail -8
