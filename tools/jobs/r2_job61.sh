b() { timeout 300 python bench.py --steps 256 --warmup 8 --no-extras --no-cpu-baseline $2 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; }
for i in 1 2; do
for r in 1 2 4; do echo "cls rows per task $r:"; LQ4_CLS_RPT=$r b x; done
done
echo "13B:"; for r in 1 2; do LQ4_CLS_RPT=$r b x "--model 13b"; done
LQ4_CLS_RPT=2 timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_ops.py -m gpu -q -x --timeout 300 -k "fp16 or logits or greedy or generate" 2>&1 | tail -2
timeout 100 python tools/trace_step.py 7b 128 8 2>&1 | grep "cls "
