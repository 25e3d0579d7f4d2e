mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/r1_tests17.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r1_tests17.log
timeout 300 python tools/trace_step.py 7b 128 6 2>&1 | tail -18
