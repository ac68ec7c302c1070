timeout 900 ncu --set full --clock-control none --import-source on -k regex:'range_(apply_pc|stats_pc)_kernel' -s 4 -c 2 -o gpurun_out/r1i_k2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1i_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 200 python tools/time_apply.py 2>&1 | grep -v "^$" | tail -1
RANGE_PC_COOP=1 timeout 200 python tools/time_apply.py 2>&1 | grep -v "^$" | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1i_bench.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/r1i_bench.json')); print(d['value'], d['ms_per_step'], d['config']['segments_ms'], d['roofline']['frac'], d['roofline']['stats_plus_apply']['frac'], 'e2e', d['e2e']['value'])"
