# 2 GPUs, final build: TP tests and bench N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x --timeout 600 > gpurun_out/r2_tests_tp2.log 2>&1; echo "tp tests rc=$?"; tail -3 gpurun_out/r2_tests_tp2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 256 --warmup 8 > gpurun_out/r2_bench_tp2.json 2> gpurun_out/r2_bench_tp2.err; echo "bench tp2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_tp2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['tp']['ids_match_single_gpu'], d['single_gpu']['value'], d['replicas']['value'])"
