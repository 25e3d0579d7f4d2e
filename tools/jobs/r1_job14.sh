mkdir -p gpurun_out
timeout 120 python tools/prof_ffn.py 4096 11008 64
timeout 600 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/r1_tests14.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r1_tests14.log
timeout 300 python tools/trace_step.py 7b 128 2>&1 | tail -9
timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r1_bench14.json 2> gpurun_out/r1_bench14.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r1_bench14.json; tail -5 gpurun_out/r1_bench14.err
