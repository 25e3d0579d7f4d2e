mkdir -p gpurun_out
LQ4_TP_REPL_O=1 timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for v in 1 0; do
LQ4_TP_REPL_O=$v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus 2 --steps 256 --warmup 8 > gpurun_out/tp2_repl$v.json 2> gpurun_out/tp2_repl$v.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/tp2_repl$v.json').read().strip().splitlines()[-1]); print('repl_o=$v', d['value'], d['ms_per_step'], d['tp']['ids_match_single_gpu'], d['single_gpu']['value'])"
done
