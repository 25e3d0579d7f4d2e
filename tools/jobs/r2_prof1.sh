# ncu evidence, round 2: tensor-core GEMM (full set, one launch), decode step (full set), classifier trace
mkdir -p gpurun_out
timeout 120 python tools/prof_gemm.py 16384 4096 11008 3 2>&1 | tail -3
timeout 120 python tools/prof_gemm.py 16384 4096 4096 3 2>&1 | tail -3
timeout 120 python tools/prof_gemm.py 16384 11008 4096 3 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_q4_tc -s 1 -c 1 -f -o gpurun_out/r2_prof_gemm python tools/prof_gemm.py 16384 4096 11008 2 > gpurun_out/r2_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 120 python tools/trace_step.py 7b 128 160 2>&1 | grep -v "^  layer\|Loading\|^dim\|^hidden\|^n_\|^seq\|^vocab\|^rope\|^Model\|^$\|slowest\|CTAs with" > gpurun_out/r2_trace_cls.txt; cat gpurun_out/r2_trace_cls.txt | tail -24
