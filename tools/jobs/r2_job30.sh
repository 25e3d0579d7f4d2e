for i in 1 2; do timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x --timeout 600 -k "dead_peer or TINY-2 or refuses" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 300 2>&1 | tail -2
