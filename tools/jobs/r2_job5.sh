mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_e2e.py -m gpu -q -x --timeout 300 -k "long_context" 2>&1 | tail -3
for W in 8 9 10 11; do echo "NWC=$W"; LQ4_NWC=$W timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
LQ4_NWC=8 timeout 120 python tools/trace_step.py 7b 128 8 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$\|slowest\|CTAs with" | head -16
