# parity tests + bench (both arms) on one B200:   gpurun --timeout 2400 -- 'bash tools/jobs/gpu_check.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 90 > gpurun_out/tests_gpu.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 256 --warmup 8 > gpurun_out/bench_7b.json 2> gpurun_out/bench_7b.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_7b.json; tail -3 gpurun_out/bench_7b.err
timeout 900 python bench.py --impl reference --steps 256 --warmup 8 > gpurun_out/bench_7b_ref.json 2> gpurun_out/bench_7b_ref.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_7b_ref.json
