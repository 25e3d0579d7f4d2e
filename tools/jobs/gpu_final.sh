mkdir -p gpurun_out
bash tools/jobs/gpu_check.sh
timeout 600 python bench.py --model 13b --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/bench_13b.json 2> gpurun_out/bench_13b.err; echo "13b rc=$?"; grep -o '"value": [0-9.]*, "unit": "tokens/s", "n_gpus"' gpurun_out/bench_13b.json
bash tools/jobs/gpu_profile.sh > gpurun_out/profile_job.out 2>&1; tail -5 gpurun_out/profile_job.out
