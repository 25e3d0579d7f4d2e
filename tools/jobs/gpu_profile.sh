# ncu evidence for profiles/:   gpurun --timeout 2400 -- 'bash tools/jobs/gpu_profile.sh'
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 6 -c 1 -f -o gpurun_out/prof_step python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu step rc=$?"
timeout 300 python tools/trace_step.py 7b 128 8 2>&1 | tail -30
