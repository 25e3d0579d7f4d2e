mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2_tests14.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests14.log
timeout 600 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r2_bench14.json 2>/dev/null; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench14.json | head -1
for pos in 128 1024; do timeout 120 python tools/trace_step.py 7b $pos 6 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$\|slowest\|CTAs with\|x staged\|raw x\|rms scale\|pairs staged\|meta landed\|first task\|ring chunks\|duration x" ; done > gpurun_out/r2_trace_attn14.txt 2>&1; cat gpurun_out/r2_trace_attn14.txt
