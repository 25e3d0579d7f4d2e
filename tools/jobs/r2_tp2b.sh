# 2 GPUs: TP tests, memcheck over a TP=2 tiny decode, bench N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x --timeout 600 > gpurun_out/r2_tests_tp2.log 2>&1; echo "tp tests rc=$?"; tail -3 gpurun_out/r2_tests_tp2.log
python - <<'PY'
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import helpers as H, llama_cu_awq_b200 as E
lib = E.lib(); c = E.Config(**H.TINY)
assert lib.lq4_write_synth_model(b"/tmp/tiny_tp.bin", C.byref(c), 11) > 0
PY
timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/r2_memcheck_tp2.%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29544 tests/tp_worker.py /tmp/tiny_tp.bin 12 /tmp/tp_out.json 1,35,72 > gpurun_out/r2_memcheck_tp2.out 2>&1; echo "memcheck tp2 rc=$?"; cat /tmp/tp_out.json | cut -c1-200; for f in gpurun_out/r2_memcheck_tp2.*.log; do tail -1 $f; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 256 --warmup 8 > gpurun_out/r2_bench_tp2.json 2> gpurun_out/r2_bench_tp2.err; echo "bench tp2 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_tp2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['tp']['ids_match_single_gpu'], d['single_gpu']['value'])"
