# round 2, first GPU call: parity tests, both bench arms, per-op traces, sanitizer, one A/B build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests1.log
timeout 900 python bench.py --steps 256 --warmup 8 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench1.json; tail -3 gpurun_out/r2_bench1.err
timeout 900 python bench.py --impl reference --steps 256 --warmup 8 > gpurun_out/r2_bench1_ref.json 2> gpurun_out/r2_bench1_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_bench1_ref.json; tail -3 gpurun_out/r2_bench1_ref.err
for o in 5 6 7 8 9; do timeout 120 python tools/trace_step.py 7b 128 $o 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$"; done > gpurun_out/r2_trace1.txt 2>&1
timeout 120 python tools/trace_step.py 7b 1024 6 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$" >> gpurun_out/r2_trace1.txt 2>&1
head -50 gpurun_out/r2_trace1.txt
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_memcheck.log python tools/sanitize_tiny.py 6 > gpurun_out/r2_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_memcheck.log
if [ -f build/lib_ffma.so ]; then LQ4_LIB=build/lib_ffma.so timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1_ffma.json 2>/dev/null; echo "ffma rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench1_ffma.json | head -1; fi
