mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 5 -c 1 -f -o gpurun_out/r1_prof_ffn_v3 python tools/prof_ffn.py 4096 11008 4 > gpurun_out/r1_prof_ffn_v3.log 2>&1
echo "ncu ffn rc=$?"; tail -3 gpurun_out/r1_prof_ffn_v3.log
