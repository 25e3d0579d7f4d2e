mkdir -p gpurun_out
for N in 8 4; do
LQ4_TP_REPL_O=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 256 --warmup 8 > gpurun_out/tp${N}_repl1.json 2> gpurun_out/tp${N}_repl1.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/tp${N}_repl1.json').read().strip().splitlines()[-1]); print('N=$N repl_o=1', d['value'], d['ms_per_step'], d['tp']['ids_match_single_gpu'], d['single_gpu']['value'])"
done
