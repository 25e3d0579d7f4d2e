# tests + bench (both arms) + ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/r1_tests6.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r1_tests6.log
timeout 900 python bench.py --steps 256 --warmup 8 > gpurun_out/r1_bench6.json 2> gpurun_out/r1_bench6.err
echo "bench rc=$?"; cat gpurun_out/r1_bench6.json; tail -5 gpurun_out/r1_bench6.err
timeout 900 python bench.py --impl reference --steps 256 --warmup 8 > gpurun_out/r1_bench6_ref.json 2> gpurun_out/r1_bench6_ref.err
echo "ref rc=$?"; cat gpurun_out/r1_bench6_ref.json; tail -5 gpurun_out/r1_bench6_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches6.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r1_ncu_bench6.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r1_ncu_bench6.log; wc -l gpurun_out/r1_launches6.csv
