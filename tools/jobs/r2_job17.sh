mkdir -p gpurun_out
for A in 0 24 48 0 24; do echo "prefetch ahead $A"; LQ4_LIB=build/lib_pf$A.so timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
LQ4_LIB=build/lib_pf24.so timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 300 2>&1 | tail -2
LQ4_LIB=build/lib_pf24.so timeout 120 python tools/trace_step.py 7b 128 9 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$\|slowest\|CTAs with" | head -24
LQ4_LIB=build/lib_pf24.so LQ4_NOMATH=1 timeout 120 python tools/trace_step.py 7b 128 9 2>&1 | grep "us each"
LQ4_LIB=build/lib_pf0.so LQ4_NOMATH=1 timeout 120 python tools/trace_step.py 7b 128 9 2>&1 | grep "us each"
