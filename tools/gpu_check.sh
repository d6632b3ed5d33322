# Quick GPU check of a kernel change (run through gpurun): small forwards against the oracle, the segmentation
# test files, a short bench.  tools/gpu_check.sh TAG [trace]
TAG=${1:-chk}
( timeout 120 python tools/quick_seg.py 1 24 && timeout 120 python tools/quick_seg.py 3 300 && UPTO=5 timeout 120 python tools/quick_seg.py 2 129 ) > gpurun_out/${TAG}_quick.txt 2>&1 || { tail -5 gpurun_out/${TAG}_quick.txt; exit 1; }
cat gpurun_out/${TAG}_quick.txt
timeout 900 python -m pytest tests/test_gpu_segment.py tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/${TAG}_seg_tests.log 2>&1; tail -3 gpurun_out/${TAG}_seg_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python tools/show_bench.py gpurun_out/${TAG}_bench.json
if [ "$2" = "trace" ]; then timeout 300 python tools/trace_gcn.py > gpurun_out/${TAG}_gcn_trace.txt 2>&1; fi
