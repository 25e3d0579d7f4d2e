# 8-GPU box: tensor-parallel parity at 4 and 8 ranks (7B), then the bench's default N=8 and N=4 paths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x --timeout 600 -k "7B or SMALL-8" > gpurun_out/r2_tests_tp8.log 2>&1; echo "tp tests rc=$?"; tail -4 gpurun_out/r2_tests_tp8.log
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 256 --warmup 8 > gpurun_out/r2_bench_tp$N.json 2> gpurun_out/r2_bench_tp$N.err; echo "bench tp$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_tp$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "tp", d["tp"], "single", d["single_gpu"]["value"], "replicas", d["replicas"]["value"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2_bench_tp$N.err").read()[-1500:])
PY
done
