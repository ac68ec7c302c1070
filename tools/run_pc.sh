timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
RANGE_APPLY_KERNEL=pc timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1e_bench.json 2> gpurun_out/r1e_bench.err; cat gpurun_out/r1e_bench.json
