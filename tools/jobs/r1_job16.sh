mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 60 > gpurun_out/r1_tests16.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r1_tests16.log
timeout 300 python tools/trace_step.py 7b 128 6 2>&1 | tail -18
LQ4_SLOT_BYTES=11264 timeout 300 python tools/trace_step.py 7b 128 9 2>&1 | tail -18
