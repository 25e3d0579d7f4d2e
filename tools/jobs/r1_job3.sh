mkdir -p gpurun_out
./build/ubench_pipes > gpurun_out/r1_ubench_pipes.txt 2>&1
cat gpurun_out/r1_ubench_pipes.txt
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "matmul_fp16" > gpurun_out/r1_sanitizer_fp16.log 2>&1
grep -n "Invalid\|at \|by thread\|Address\|ERROR SUMMARY" gpurun_out/r1_sanitizer_fp16.log | head -30
