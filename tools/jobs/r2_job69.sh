timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 300 2>&1 | tail -3
b() { LQ4_LIB=$1 timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
for i in 1 2; do
echo "pv split:"; b build/lib_PV.so
echo "base:"; b build/lib_base2.so
done
timeout 100 python tools/trace_step.py 7b 128 6 2>&1 | grep "step at\|attn \|first weights\|warp0 done\|all warps done"
