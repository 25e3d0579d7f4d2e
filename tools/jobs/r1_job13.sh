mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 5 -c 1 -f -o gpurun_out/r1_prof_ffn_v4 python tools/prof_ffn.py 4096 11008 4 > gpurun_out/r1_prof_ffn_v4.log 2>&1
echo "ncu ffn rc=$?"; tail -2 gpurun_out/r1_prof_ffn_v4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 5 -c 1 -f -o gpurun_out/r1_prof_step_v4 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1_prof_step_v4.log 2>&1
echo "ncu step rc=$?"; tail -2 gpurun_out/r1_prof_step_v4.log | cut -c1-300
