b() { LQ4_LIB=$1 timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
echo "rotated, 1 copy per ffn slot (timing only):"; b build/lib_FFN1.so
echo "rotated:"; b build/lib_ROT.so
echo "base:"; b build/lib_base.so
LQ4_LIB=build/lib_FFN1.so timeout 150 python tools/trace_step.py 7b 128 8 2>&1 | grep "step at\|ffn \|gemv \|x staged  \|all warps done  "
