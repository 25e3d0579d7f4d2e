mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -q --timeout 200 > gpurun_out/r1_tests5.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/r1_tests5.log
timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r1_bench5.json 2> gpurun_out/r1_bench5.err
echo "bench rc=$?"; cat gpurun_out/r1_bench5.json; tail -5 gpurun_out/r1_bench5.err
