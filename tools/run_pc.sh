RANGE_APPLY_KERNEL=pc timeout 120 python tools/check_pc.py 2>&1 | tail -2
timeout 200 python tools/time_apply.py 2>&1 | grep -v "^$" | tail -1
PROF=1 RANGE_APPLY_KERNEL=pc timeout 200 python tools/time_apply.py 2>&1 | grep -v "^$" | tail -8
