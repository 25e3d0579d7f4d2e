mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r1_tests8.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r1_tests8.log
timeout 300 python tools/prof_ffn.py 4096 11008 64
timeout 900 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r1_bench8.json 2> gpurun_out/r1_bench8.err
echo "bench rc=$?"; cat gpurun_out/r1_bench8.json; tail -5 gpurun_out/r1_bench8.err
