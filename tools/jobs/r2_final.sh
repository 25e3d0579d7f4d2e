# end-of-round evidence: full GPU parity suite, smoke, both bench arms, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_final_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["model_13b"]["value"], d["prefill"]["value"], d["prefill"]["roofline"]["achieved"], d["fused0"]["value"], d["clocks"])
r=json.loads(open("gpurun_out/r2_bench_final_ref.json").read().strip().splitlines()[-1])
print(r["value"], r["ms_per_step"])
PY
