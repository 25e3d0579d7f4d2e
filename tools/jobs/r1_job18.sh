mkdir -p gpurun_out
timeout 900 python bench.py --steps 256 --warmup 8 > gpurun_out/r1_bench18.json 2> gpurun_out/r1_bench18.err
echo "bench rc=$?"; cat gpurun_out/r1_bench18.json | cut -c1-2500; tail -3 gpurun_out/r1_bench18.err
timeout 900 python bench.py --model 13b --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/r1_bench18_13b.json 2> gpurun_out/r1_bench18_13b.err
echo "bench13 rc=$?"; cat gpurun_out/r1_bench18_13b.json | cut -c1-700; tail -3 gpurun_out/r1_bench18_13b.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1_launches18.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench18.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:interp_kernel -s 6 -c 1 -f -o gpurun_out/r1_prof_step_v5 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r1_prof_step_v5.log 2>&1
echo "ncu step rc=$?"
