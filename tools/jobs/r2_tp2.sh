# tensor-parallel check on 2 GPUs: parity tests + the bench's default N=2 path
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x --timeout 600 > gpurun_out/r2_tests_tp2.log 2>&1; echo "tp tests rc=$?"; tail -4 gpurun_out/r2_tests_tp2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 256 --warmup 8 > gpurun_out/r2_bench_tp2.json 2> gpurun_out/r2_bench_tp2.err; echo "bench tp2 rc=$?"; tail -c 1500 gpurun_out/r2_bench_tp2.json; tail -5 gpurun_out/r2_bench_tp2.err
