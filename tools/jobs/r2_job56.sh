timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_ops.py -m gpu -q -x --timeout 300 2>&1 | tail -3
b() { LQ4_LIB=$1 timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
for i in 1 2; do
echo "dynamic tasks:"; b build/lib_DYN.so
echo "base:"; b build/lib_base.so
done
timeout 150 python tools/trace_step.py 7b 128 8 2>&1 | grep -A14 "step at\|per warp, us" | grep -v "layer 1"
