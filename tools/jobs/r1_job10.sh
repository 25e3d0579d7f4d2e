mkdir -p gpurun_out
timeout 300 python tools/prof_ffn.py 4096 11008 64
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r1_tests11.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r1_tests11.log
timeout 900 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r1_bench11.json 2> gpurun_out/r1_bench10.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r1_bench11.json; tail -5 gpurun_out/r1_bench10.err
