timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/time_apply.py 2>&1 | grep -v "^$" | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['segments_ms'], d['roofline']['frac'], 'e2e', d['e2e']['value'])"
