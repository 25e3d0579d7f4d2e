mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/r1_tests15.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r1_tests15.log
timeout 300 python tools/trace_step.py 7b 128 8 2>&1 | tail -18
timeout 300 python tools/trace_step.py 7b 128 9 2>&1 | tail -9
