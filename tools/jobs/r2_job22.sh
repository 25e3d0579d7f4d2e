mkdir -p gpurun_out
for i in 1 2; do timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
timeout 300 python bench.py --steps 1024 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
timeout 300 python bench.py --model 13b --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2_tests22.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests22.log
