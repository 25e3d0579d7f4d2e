mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 5 -c 1 -f -o gpurun_out/r1_prof_ffn python tools/prof_ffn.py 4096 11008 4 > gpurun_out/r1_prof_ffn.log 2>&1
echo "ncu ffn rc=$?"; tail -3 gpurun_out/r1_prof_ffn.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 5 -c 1 -f -o gpurun_out/r1_prof_step python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1_prof_step.log 2>&1
echo "ncu step rc=$?"; tail -3 gpurun_out/r1_prof_step.log
ls -la gpurun_out/*.ncu-rep
