mkdir -p gpurun_out
timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
timeout 300 python bench.py --steps 1024 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
LQ4_LIB=build/lib_A.so timeout 300 python bench.py --steps 1024 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2_tests20.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests20.log
for pos in 128 1024; do timeout 120 python tools/trace_step.py 7b $pos 6 2>&1 | grep "step at\|attn \|all warps done\|first weights\|warp0 done"; done
