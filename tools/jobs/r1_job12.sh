mkdir -p gpurun_out
timeout 120 python tools/prof_ffn.py 4096 11008 64
timeout 600 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/r1_tests12.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r1_tests12.log
timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r1_bench12.json 2> gpurun_out/r1_bench12.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r1_bench12.json; tail -5 gpurun_out/r1_bench12.err
