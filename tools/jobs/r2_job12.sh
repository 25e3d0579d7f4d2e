mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prefill.py -m gpu -q -x --timeout 200 2>&1 | tail -3
for s in "16384 4096 11008" "16384 4096 4096" "16384 11008 4096"; do timeout 120 python tools/prof_gemm.py $s 3 2>&1 | tail -1; done
echo "--- opk: stand-alone kernels + PDL behind the operator API"
LQ4_OPK=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -q -x --timeout 300 2>&1 | tail -3
for o in 0 1; do LQ4_OPK=$o timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/r2_bench_opk$o.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_opk$o.json').read().strip().splitlines()[-1]); print('opk=$o fused0', d['fused0']['value'], d['fused0']['ms_per_step'], d['fused0']['ids_equal_fused'], 'gemv op us', d['gemv_4096_op']['us_per_launch'], 'ffn op us', d['roofline_ffn_op']['us_per_launch'], 'prefill', d['prefill']['value'], d['prefill']['ms_gemm'], d['prefill']['roofline']['achieved'])"; done
