mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_ops.py -m gpu -q -x --timeout 300 2>&1 | tail -3
b() { LQ4_LIB=$1 timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
for i in 1 2; do
echo "rotated start + rebalance:"; b build/lib_ROT.so
echo "base:"; b build/lib_base.so
done
timeout 120 python tools/trace_step.py 7b 128 9 2>&1 | grep -A12 "step at\|^op \|per warp" | grep -v "slowest\|CTAs with\|duration\|SM-clock\|raw x\|rms\|pairs\|meta\|first task\|arrive ->\|layer 1"
