nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
SORT=1 PROF=1 python tools/time_apply.py 2>&1 | tail -8
SORT=0 PROF=1 python tools/time_apply.py 2>&1 | tail -8
