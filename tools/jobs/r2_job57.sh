b() { LQ4_LIB=$1 timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
for i in 1 2; do
echo "base, 1 copy per ffn slot (timing only):"; b build/lib_base_FFN1.so
echo "dynamic, 1 copy per ffn slot (timing only):"; b build/lib_DYN_FFN1.so
echo "base:"; b build/lib_base.so
done
