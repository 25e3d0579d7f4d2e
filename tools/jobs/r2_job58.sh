timeout 900 python -m pytest tests/test_gpu_prefill.py -m gpu -q -x --timeout 600 2>&1 | tail -4
timeout 600 python bench.py --steps 64 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        d = json.loads(l); print(json.dumps(d.get('prefill'), indent=1)[:900]); print('ms_per_step', d['ms_per_step'])
"
