mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/r1_tests20.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r1_tests20.log
timeout 300 python tools/trace_step.py 7b 128 8 2>&1 | grep -A8 "step at"
timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r1_bench20.json 2> gpurun_out/r1_bench20.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/r1_bench20.json; tail -3 gpurun_out/r1_bench20.err
