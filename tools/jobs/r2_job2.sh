# round 2, job 2: A/B of consumer-warp counts (11 / 13 / 15) + the early ring wait
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_ops.py -m gpu -q -x --timeout 300 > gpurun_out/r2_tests2.log 2>&1; echo "tests main rc=$?"; tail -2 gpurun_out/r2_tests2.log
timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r2_bench2_main.json 2>/dev/null; echo "main rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench2_main.json | head -1
for W in 13 15; do
  LQ4_LIB=build/lib_w$W.so timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 300 > gpurun_out/r2_tests2_w$W.log 2>&1; echo "tests w$W rc=$?"; tail -2 gpurun_out/r2_tests2_w$W.log
  LQ4_LIB=build/lib_w$W.so timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r2_bench2_w$W.json 2>/dev/null; echo "w$W rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench2_w$W.json | head -1
  LQ4_LIB=build/lib_w$W.so timeout 120 python tools/trace_step.py 7b 128 8 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$\|slowest\|CTAs with" > gpurun_out/r2_trace2_w$W.txt; head -16 gpurun_out/r2_trace2_w$W.txt
done
timeout 120 python tools/trace_step.py 7b 128 8 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$\|slowest\|CTAs with" > gpurun_out/r2_trace2_main.txt; head -30 gpurun_out/r2_trace2_main.txt
