timeout 900 python -m pytest tests/test_gpu_prefill.py -m gpu -q -x --timeout 600 2>&1 | tail -3
bash tools/jobs/r2_job59.sh 2>&1 | tail -12
timeout 300 python tools/prof_prefill.py 8 2048 2 2>&1 | tail -2
