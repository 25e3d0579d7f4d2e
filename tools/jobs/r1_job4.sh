mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 120 > gpurun_out/r1_ops.log 2>&1
echo "ops rc=$?"; tail -15 gpurun_out/r1_ops.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q --timeout 200 > gpurun_out/r1_e2e.log 2>&1
echo "e2e rc=$?"; tail -15 gpurun_out/r1_e2e.log
