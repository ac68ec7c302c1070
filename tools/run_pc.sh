python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
N=6144 M=50000 timeout 120 python tools/check_pc.py 2>&1 | tail -2
N=7900 M=50000 timeout 120 python tools/check_pc.py 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['segments_ms'], 'e2e', d['e2e']['value'], d['gpu_launches'])"
timeout 300 python tools/time_e2e.py 2>&1 | tail -4
